"""CPU checks of the fp16-storage precision model (oracle/halfprec.py) and of the re-balanced / amplified test weights:
the model stays within fp16-storage distance of the fp32 oracle, and the feature comparison the GPU tests use is sensitive to
a single deep conv (what the default-init 8-bit comparison is not)."""
import copy

import numpy as np
import torch

from util import inrange_state_dict, max_lsb, oracle_model_from_sd, psnr_u8, random_state_dict, rel_l2, synth_frame


def _x(f):
    return torch.from_numpy(np.ascontiguousarray(f[:, :, ::-1].astype(np.float32) / 255.0)).permute(2, 0, 1)[None]


def _img(o):
    return (o[0].clamp(0, 1).permute(1, 2, 0).numpy()[:, :, ::-1] * 255.0).round().astype(np.uint8)


def test_fp16_storage_model_tracks_fp32_oracle():
    from oracle import halfprec as HP

    for name in ("RealESRGAN_x4plus_anime_6B", "RealESRGAN_x2plus", "RealESRGAN_x4_v3"):
        sd = inrange_state_dict(name, 0, gain=10.0)
        m = oracle_model_from_sd(name, sd)
        f = synth_frame(32, 40, seed=3)
        with torch.no_grad():
            fe = {}
            o16 = HP.forward_fp16_storage(m, _x(f), fe)
            o32 = m(_x(f))
        a, b = _img(o16), _img(o32)
        assert max_lsb(a, b) <= 1 and psnr_u8(a, b) >= 50.0, name
        assert ((b == 0) | (b == 255)).mean() < 0.01 and b.std() > 2.0   # in-range weights: nothing clamps
        if name != "RealESRGAN_x4_v3":
            f32 = HP.rrdbnet_features_fp32(m, _x(f))
            for k in ("feat", "body", "trunk"):
                assert fe[k].shape == f32[k].shape and rel_l2(fe[k], f32[k]) < 3e-3, (name, k)


def test_default_init_hides_the_body_and_clamps():
    """Why the extra weight variants exist: measured facts about the default random init."""
    from oracle.pipeline import OracleRestorer

    name = "RealESRGAN_x4plus_anime_6B"
    f = synth_frame(32, 40, seed=3)
    sd = random_state_dict(name, 0)
    m = oracle_model_from_sd(name, sd)
    m2 = copy.deepcopy(m)
    with torch.no_grad():
        m2.body[3].rdb2.conv3.weight *= 1.05
        a = m.body(m.conv_first(_x(f)))
        b = m2.body(m2.conv_first(_x(f)))
    assert rel_l2(b.numpy(), a.numpy()) < 5e-5    # a 5 % error in a deep conv is invisible next to fp16 noise (1e-3)
    o = OracleRestorer(name, tile=1024, tile_pad=10, model=m)
    out = o.process_frame(f)
    assert out.std() > 2.0


def test_amplified_weights_make_a_deep_conv_visible():
    from oracle import halfprec as HP

    name = "RealESRGAN_x4plus_anime_6B"
    sd = inrange_state_dict(name, 0, gain=10.0)
    m = oracle_model_from_sd(name, sd)
    m2 = copy.deepcopy(m)
    with torch.no_grad():
        m2.body[3].rdb2.conv3.weight *= 1.05
    f = synth_frame(32, 40, seed=3)
    fa, fb = {}, {}
    HP.rrdbnet_fp16_storage(m, _x(f), fa)
    HP.rrdbnet_fp16_storage(m2, _x(f), fb)
    assert rel_l2(fb["body"], fa["body"]) > 8e-4   # twice the GPU tests' 4e-4 tolerance, from ONE conv 5 % off
