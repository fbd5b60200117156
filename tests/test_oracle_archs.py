"""Oracle networks: structural cross-checks against the published upstream models.
The upstream packages are absent, so these pin STRUCTURE (parameter counts, key sets, shapes), not outputs."""
import numpy as np
import pytest
import torch

from oracle.archs import MODEL_ZOO, build_model, pixel_unshuffle
from video_restore_b200.models import MODEL_ZOO as PZOO
from video_restore_b200.models import conv_layers, flops_per_input_pixel

PUBLISHED_PARAMS = {"RealESRGAN_x4plus": 16_697_987, "RealESRGAN_x2plus": 16_703_171,
                    "RealESRGAN_x4plus_anime_6B": 4_467_779, "RealESRGAN_x4_v3": 1_213_296}
PUBLISHED_TENSORS = {"RealESRGAN_x4plus": 702, "RealESRGAN_x2plus": 702, "RealESRGAN_x4plus_anime_6B": 192,
                     "RealESRGAN_x4_v3": 101}
FLOPS = {"RealESRGAN_x4plus": 35_853_696, "RealESRGAN_x2plus": 8_966_016,
         "RealESRGAN_x4plus_anime_6B": 11_412_864, "RealESRGAN_x4_v3": 2_418_048}


@pytest.mark.parametrize("name", sorted(MODEL_ZOO))
def test_parameter_counts_and_keys(name):
    m = build_model(name, seed=0)
    assert sum(p.numel() for p in m.parameters()) == PUBLISHED_PARAMS[name]
    sd = m.state_dict()
    assert len(sd) == PUBLISHED_TENSORS[name]
    # the product's layer list names exactly the conv tensors of the oracle's state_dict
    for prefix, cin, cout, _ in conv_layers(PZOO[name]):
        assert tuple(sd[prefix + ".weight"].shape) == (cout, cin, 3, 3)
        assert tuple(sd[prefix + ".bias"].shape) == (cout,)
    assert flops_per_input_pixel(PZOO[name]) == FLOPS[name]


def test_upstream_key_names():
    sd = build_model("RealESRGAN_x4plus", 0).state_dict()
    for k in ("conv_first.weight", "body.0.rdb1.conv1.weight", "body.22.rdb3.conv5.bias", "conv_body.weight",
              "conv_up1.weight", "conv_up2.weight", "conv_hr.weight", "conv_last.bias"):
        assert k in sd
    sv = build_model("RealESRGAN_x4_v3", 0).state_dict()
    assert "body.0.weight" in sv and "body.1.weight" in sv and "body.66.bias" in sv and "body.65.weight" in sv
    assert tuple(sv["body.66.weight"].shape) == (48, 64, 3, 3) and tuple(sv["body.1.weight"].shape) == (64,)


def test_pixel_unshuffle_channel_order():
    x = torch.arange(2 * 3 * 4 * 6, dtype=torch.float32).view(2, 3, 4, 6)
    y = pixel_unshuffle(x, 2)
    assert y.shape == (2, 12, 2, 3)
    for c in range(3):
        for dy in range(2):
            for dx in range(2):
                assert torch.equal(y[:, c * 4 + dy * 2 + dx], x[:, c, dy::2, dx::2])
    assert torch.equal(torch.nn.functional.pixel_unshuffle(x, 2), y)  # torch's own op agrees
    with pytest.raises(RuntimeError):
        pixel_unshuffle(torch.zeros(1, 3, 5, 6), 2)


def test_output_shapes_and_x2_modpad():
    from oracle.realesrganer import RealESRGANer

    f = np.random.default_rng(0).integers(0, 256, (17, 13, 3), dtype=np.uint8)
    for name in ("RealESRGAN_x4_v3", "RealESRGAN_x4plus_anime_6B"):
        out, mode = RealESRGANer(4, build_model(name, 0), tile=16, tile_pad=2, pre_pad=0).enhance(f, outscale=4)
        assert out.shape == (68, 52, 3) and out.dtype == np.uint8 and mode == "RGB"
    # outscale != scale: Lanczos resize of the network-scale result (upstream's last step)
    import cv2
    up = RealESRGANer(4, build_model("RealESRGAN_x4_v3", 0), tile=16, tile_pad=2, pre_pad=0)
    full, _ = up.enhance(f, outscale=4)
    half, _ = up.enhance(f, outscale=2)
    assert half.shape == (34, 26, 3)
    assert np.array_equal(half, cv2.resize(full, (26, 34), interpolation=cv2.INTER_LANCZOS4))
    m2 = build_model("RealESRGAN_x2plus", 0)
    out, _ = RealESRGANer(2, m2, tile=8, tile_pad=2, pre_pad=0).enhance(f, outscale=2)  # 17x13 -> mod-pad 18x14
    assert out.shape == (34, 26, 3)


def test_rdb_init_scale():
    m = build_model("RealESRGAN_x4plus_anime_6B", 0)
    w_rdb = m.body[0].rdb1.conv1.weight
    w_plain = m.conv_body.weight
    assert float(m.body[0].rdb1.conv1.bias.abs().max()) == 0.0
    # kaiming_normal * 0.1 : std = 0.1 * sqrt(2 / (64*9))
    assert abs(float(w_rdb.std()) - 0.1 * np.sqrt(2.0 / 576)) < 1e-3
    assert float(w_plain.abs().max()) <= 1.0 / np.sqrt(576) + 1e-6  # torch Conv2d default: U(-1/sqrt(fan_in), ..)
