"""Parity AT SIZE on every BASELINE.json config, and network-level checks that are sensitive to the body (`-m gpu`).

North star: "uint8 output frames within +-1 LSB (PSNR >= 50 dB against the fp32 reference) ... with the numerical tolerance
met on every config". The CPU oracle (reference semantics: video_upscaler.py:490-505 with the preset table :687-701) is run
ONCE on the FULL frame of configs 3, 4 (both presets) and 5 -- about 15-60 s each on the GPU box's host cores -- and compared
with the CUDA path over the whole frame. (Round 1 only compared CUDA with CUDA at these sizes.)

The second half answers "would a wrong layer deep in the body be seen?": with default-init weights the trunk is 66 x
conv_first's output and 42 % of the x4plus frame clamps, so the tests below use re-balanced (`inrange_state_dict`) and
amplified weights and compare FEATURE tensors (vr_debug_activation) against the fp32 oracle and against its fp16-storage
emulation (oracle/halfprec.py): a dropped conv or two swapped blocks anywhere in the 23-block body violate the tolerance
10 x over, and on 1-3 block models one conv being 5 % off does.
"""
import os
import time

import numpy as np
import pytest

from util import (inrange_state_dict, max_lsb, oracle_model_from_sd, psnr_u8, psnr_unsaturated, random_state_dict, rel_l2,
                  synth_frame)

pytestmark = pytest.mark.gpu

LSB_TOL, PSNR_TOL = 1, 50.0


def _oracle(name, sd, tile, pad, blend):
    import torch

    from oracle.pipeline import OracleRestorer

    torch.set_num_threads(os.cpu_count() or 1)
    return OracleRestorer(name, tile=tile, tile_pad=pad, blend=blend, model=oracle_model_from_sd(name, sd))


def _gpu(name, sd, tile, pad, blend):
    from video_restore_b200.restorer import FrameRestorer

    return FrameRestorer(name, sd, tile=tile, tile_pad=pad, blend=blend)


def _report(tag, out, ref, t_cpu, lsb_tol=LSB_TOL, over1_frac=0.0):
    lsb, p = max_lsb(out, ref), psnr_u8(out, ref)
    pu, frac = psnr_unsaturated(out, ref)
    d = np.abs(out.astype(np.int16) - ref.astype(np.int16))
    n_over = int((d > 1).sum())
    print(f"\n[fullsize] {tag}: {ref.shape[1]}x{ref.shape[0]} max {lsb} LSB ({n_over} of {d.size} values off by more than 1), "
          f"PSNR {p:.2f} dB, PSNR over the {frac:.1%} unclamped values {pu:.2f} dB, {float((d > 0).mean()):.3%} of values "
          f"differ; oracle {t_cpu:.1f} s on {os.cpu_count()} host cores")
    assert out.shape == ref.shape and out.dtype == np.uint8
    assert lsb <= lsb_tol and n_over <= over1_frac * d.size, f"{tag}: max {lsb} LSB, {n_over} values off by more than 1"
    assert p >= PSNR_TOL and pu >= PSNR_TOL, f"{tag}: PSNR {p:.2f} / {pu:.2f} dB"
    assert ref.std() > 2.0, "degenerate reference image"


# name, H, W, tile, pad, blend  -- the frame each BASELINE config / reference preset produces
C3 = ("RealESRGAN_x2plus", 1080, 1920, 512, 32, "gaussian")
FULL = {
    "c4_x4plus_720p_qmax_plain": ("RealESRGAN_x4plus", 720, 1280, 1536, 10, "crop"),        # --quality max (:690)
    "c5_x4plus_1080p": ("RealESRGAN_x4plus", 1080, 1920, 1024, 10, "crop"),
}


@pytest.mark.parametrize("cfg", sorted(FULL))
def test_full_frame_vs_oracle(gpu_lib, cfg):
    name, H, W, tile, pad, blend = FULL[cfg]
    sd = random_state_dict(name, seed=0)
    f = synth_frame(H, W, seed=13)
    gpu = _gpu(name, sd, tile, pad, blend)
    out = gpu.process_frame(f)
    gpu.close()
    orc = _oracle(name, sd, tile, pad, blend)
    t0 = time.perf_counter()
    ref = orc.process_frame(f)
    _report(cfg, out, ref, time.perf_counter() - t0)


def test_full_frame_c3_x2plus_vs_oracle(gpu_lib):
    """BASELINE configs[2]: x2plus 1080p -> 2160p, tile 512 / overlap 32, seamless Gaussian blend; 12 tiles, 24.9 M values.

    The DEFAULT random init makes x2plus an extreme case: 48 % of its frame clamps and the unclamped part spans the whole range
    (std 100 levels), i.e. the network's output gain is several times x4plus's. At that gain fp16 STORAGE itself -- what the
    reference's own half=True path does (video_upscaler.py:335,714) -- is occasionally 2 levels away from fp32: measured on
    B200 in round 2, max 2 LSB at PSNR 59.9 dB. The test therefore pins three things: (1) against the fp32 oracle: PSNR >= 50 dB
    and at most one value in a million off by more than 1 level, none by more than 2; (2) against the oracle evaluated with
    fp16 storage (oracle/halfprec.py), i.e. the same precision model on the CPU: +-1 LSB everywhere -- the 2-level events
    belong to the precision the reference prescribes, not to the kernels; (3) with re-balanced weights whose output stays in
    range (tests/util.py::inrange_state_dict): strictly +-1 LSB / PSNR >= 50 dB against fp32."""
    from oracle.halfprec import HalfStorageNet
    from oracle.pipeline import OracleRestorer

    name, H, W, tile, pad, blend = C3
    f = synth_frame(H, W, seed=13)
    sd = random_state_dict(name, seed=0)
    gpu = _gpu(name, sd, tile, pad, blend)
    out = gpu.process_frame(f)
    gpu.close()
    orc = _oracle(name, sd, tile, pad, blend)
    t0 = time.perf_counter()
    ref = orc.process_frame(f)
    _report("c3_x2plus_1080p_seamless vs fp32 oracle", out, ref, time.perf_counter() - t0, lsb_tol=2, over1_frac=1e-6)
    o16 = OracleRestorer(name, tile=tile, tile_pad=pad, blend=blend, model=HalfStorageNet(orc.model))
    t0 = time.perf_counter()
    ref16 = o16.process_frame(f)
    _report("c3_x2plus_1080p_seamless vs fp16-storage oracle", out, ref16, time.perf_counter() - t0)
    print(f"[fullsize] c3: fp16-storage oracle vs fp32 oracle (CPU vs CPU): max {max_lsb(ref16, ref)} LSB, "
          f"PSNR {psnr_u8(ref16, ref):.2f} dB")
    sd2 = inrange_state_dict(name, seed=0)
    gpu = _gpu(name, sd2, tile, pad, blend)
    out2 = gpu.process_frame(f)
    gpu.close()
    t0 = time.perf_counter()
    ref2 = _oracle(name, sd2, tile, pad, blend).process_frame(f)
    _report("c3_x2plus_1080p_seamless, in-range weights, vs fp32 oracle", out2, ref2, time.perf_counter() - t0)
    assert float(((ref2 == 0) | (ref2 == 255)).mean()) < 0.01


def test_full_frame_c4_enhanced_vs_oracle(gpu_lib):
    """BASELINE configs[3] as `--quality max --enhanced` produces it (video_upscaler.py:690-691, :326, :495-496): tile 512,
    overlap 64 -> 6 tiles, Gaussian blend, bilateral pre-denoise; then unsharp + CLAHE + temporal. Upscale stage (bilateral ->
    tiled network -> blend) vs the oracle over the whole 5120x2880 frame; the enhancement stage bit-exactly on the GPU's own
    intermediates for two consecutive frames (CLAHE's histogram equalisation amplifies a +-1 LSB input difference on flat
    regions, so the chain's end-to-end tolerance is a PSNR figure only -- printed, asserted > 40 dB)."""
    from oracle import filters as OF
    from oracle.pipeline import FrameOpts as OOpts
    from video_restore_b200.restorer import FrameOpts

    name, H, W, tile, pad, blend = "RealESRGAN_x4plus", 720, 1280, 512, 64, "gaussian"
    sd = random_state_dict(name, seed=0)
    f0, f1 = synth_frame(H, W, seed=13, index=0), synth_frame(H, W, seed=13, index=1)
    gpu = _gpu(name, sd, tile, pad, blend)
    up0 = gpu.process_frame(f0, FrameOpts(denoise=True))
    up1 = gpu.process_frame(f1, FrameOpts(denoise=True))
    full = FrameOpts(denoise=True, sharpen=0.5, clahe=True, temporal=True)
    o0 = gpu.process_frame(f0, full)
    o1 = gpu.process_frame(f1, full)
    gpu.close()
    orc = _oracle(name, sd, tile, pad, blend)
    t0 = time.perf_counter()
    ref0 = orc.process_frame(f0, OOpts(denoise=True))
    _report("c4_x4plus_720p_qmax_enhanced (upscale stage)", up0, ref0, time.perf_counter() - t0)
    e0 = OF.clahe_bgr(OF.unsharp_mask(up0, 0.5))
    e1 = OF.clahe_bgr(OF.unsharp_mask(up1, 0.5))
    assert np.array_equal(o0, e0), "frame 0 passes through the temporal stage"
    assert np.array_equal(o1, OF.temporal_blend(e1, e0)), "unsharp -> CLAHE -> temporal at 2880p is not bit-exact"
    chain_ref = OF.clahe_bgr(OF.unsharp_mask(ref0, 0.5))
    p = psnr_u8(o0, chain_ref)
    print(f"[fullsize] c4 enhanced, whole chain vs oracle chain: PSNR {p:.2f} dB, max {max_lsb(o0, chain_ref)} LSB")
    assert p > 40.0


# ----------------------------------------------------------------------------------------------------------
# body-sensitive checks
# ----------------------------------------------------------------------------------------------------------
def _x(f):
    import torch

    return torch.from_numpy(np.ascontiguousarray(f[:, :, ::-1].astype(np.float32) / 255.0)).permute(2, 0, 1)[None]


def _oracle_net(spec, sd):
    """fp32 oracle network for a model-zoo name or an RRDBNet spec dict (shallow variants of x4plus)."""
    import torch

    from oracle.archs import RRDBNet

    if isinstance(spec, str):
        return oracle_model_from_sd(spec, sd)
    m = RRDBNet(3, 3, scale=spec["scale"], num_feat=64, num_block=spec["num_block"], num_grow_ch=32)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return m.eval()


def _features(spec, sd, f, sd_oracle=None):
    """(CUDA features, fp32 oracle features, fp16-storage features) of a single-tile frame; `sd_oracle`: other weights for the
    oracle side (to show that a difference IS detected)."""
    from oracle import halfprec as HP

    gpu = _gpu(spec, sd, 1 << 20, 10, "crop")
    gpu.process_frame(f)
    got = {k: gpu.debug_activation(k) for k in ("feat", "body", "trunk")}
    gpu.close()
    m = _oracle_net(spec, sd if sd_oracle is None else sd_oracle)
    f32 = HP.rrdbnet_features_fp32(m, _x(f))
    f16 = {}
    HP.rrdbnet_fp16_storage(m, _x(f), f16)
    return got, f32, f16


def _shallow(nb, gain):
    """x4plus cut down to its first `nb` RRDBs (spec dict + matching state_dict)."""
    sd = inrange_state_dict("RealESRGAN_x4plus", seed=0, gain=gain)
    keep = {k: v for k, v in sd.items() if not k.startswith("body.") or int(k.split(".")[1]) < nb}
    return dict(kind="rrdb", scale=4, num_block=nb, num_conv=0), keep


# tol16: rel L2 vs the fp16-storage model. With default-size dense-block weights (gain 1) the CUDA features track it to ~1e-4
# (fp32 summation order only). With gain 10 the network amplifies every rounding flip, and over 23 blocks the two fp16
# computations decorrelate to the fp16-vs-fp32 distance itself (measured 1.0e-3 for x4plus, 3.8e-4 for the 6-block model).
@pytest.mark.parametrize("name,gain,tol16", [("RealESRGAN_x4plus", 1.0, 4e-4), ("RealESRGAN_x4plus", 10.0, 3e-3),
                                             ("RealESRGAN_x2plus", 10.0, 3e-3), ("RealESRGAN_x4plus_anime_6B", 10.0, 1.5e-3)])
def test_feature_parity_amplified(gpu_lib, name, gain, tol16):
    """conv_first output, last RRDB's output and the trunk of the CUDA path vs the fp32 oracle (fp16-storage tolerance) and vs
    the fp16-storage emulation. gain 10 = kaiming-normal dense-block convs: every one of the 345 (x4plus) body convs then
    moves the features by O(1) of their magnitude, so a dropped / mis-wired layer is a 10 x tolerance violation (see
    test_feature_check_detects_a_wrong_deep_layer)."""
    sd = inrange_state_dict(name, seed=0, gain=gain)
    f = synth_frame(72, 88, seed=3)
    got, f32, f16 = _features(name, sd, f)
    for k in ("feat", "body", "trunk"):
        assert got[k].shape == f32[k].shape, (k, got[k].shape, f32[k].shape)
        e32, e16 = rel_l2(got[k], f32[k]), rel_l2(got[k], f16[k])
        print(f"\n[features] {name} gain {gain} {k}: rel L2 vs fp32 {e32:.2e}, vs fp16-storage model {e16:.2e}")
        assert e32 < 3e-3, f"{k}: {e32:.2e} vs the fp32 oracle"
        assert e16 < tol16, f"{k}: {e16:.2e} vs the fp16-storage model"
        rms = float(np.sqrt(np.mean(f16[k].astype(np.float64) ** 2)))
        assert (np.abs(got[k] - f16[k]) <= 10 * tol16 * (np.abs(f16[k]) + rms)).all(), f"{k}: element-wise outlier"


@pytest.mark.parametrize("nb", [1, 2, 3])
def test_feature_parity_shallow_blocks_detect_5_percent(gpu_lib, nb):
    """One / two / three RRDBs with kaiming-normal dense-block weights: too shallow for the roundings to decorrelate, so the
    CUDA features track the fp16-storage model to ~1e-4 -- and ONE conv of the last block being 5 % off (loaded into the CUDA
    path only) is a violation by a wide margin. Every block runs the same code and buffer rotation, so this is the per-layer
    sensitivity of the whole body."""
    spec, sd = _shallow(nb, 10.0)
    f = synth_frame(72, 88, seed=3)
    got, f32, f16 = _features(spec, sd, f)
    floor = rel_l2(got["body"], f16["body"])
    bad = dict(sd)
    key = f"body.{nb - 1}.rdb2.conv3.weight"
    bad[key] = sd[key] * np.float32(1.05)
    got_bad, _, _ = _features(spec, bad, f, sd_oracle=sd)
    seen = rel_l2(got_bad["body"], f16["body"])
    print(f"\n[features] {nb} block(s), gain 10: rel L2 vs fp16-storage model {floor:.2e} (vs fp32 {rel_l2(got['body'], f32['body']):.2e}); "
          f"with {key} 5 % off: {seen:.2e}")
    assert floor < 3e-4 and rel_l2(got["trunk"], f16["trunk"]) < 3e-4
    assert seen > 8e-4 and seen > 3 * floor


def test_feature_parity_tile_atlas(gpu_lib):
    """Same check through the tile atlas (2 x 3 tiles with gap rows / columns): every tile's region of the atlas equals the
    oracle's features of that padded tile run on its own."""
    import torch

    from oracle import halfprec as HP
    from oracle.realesrganer import tile_grid

    name, tile, pad = "RealESRGAN_x4plus_anime_6B", 48, 8
    sd = inrange_state_dict(name, seed=0, gain=10.0)
    f = synth_frame(90, 130, seed=4)
    gpu = _gpu(name, sd, tile, pad, "crop")
    gpu.process_frame(f)
    atlas = gpu.debug_activation("trunk")
    gpu.close()
    m = oracle_model_from_sd(name, sd)
    x = torch.from_numpy(np.ascontiguousarray(f[:, :, ::-1].astype(np.float32) / 255.0)).permute(2, 0, 1)[None]
    grid = tile_grid(90, 130, tile, pad, 4)
    tiles_x = -(-130 // tile)
    colw = [int(grid[j][5] - grid[j][4]) for j in range(tiles_x)]
    rowh = [int(grid[i * tiles_x][7] - grid[i * tiles_x][6]) for i in range(len(grid) // tiles_x)]
    ax = np.concatenate([[0], np.cumsum([w + 1 for w in colw])])
    ay = np.concatenate([[0], np.cumsum([h + 1 for h in rowh])])
    assert atlas.shape[:2] == (sum(rowh) + len(rowh) - 1, sum(colw) + len(colw) - 1)
    for ti, t in enumerate(grid.tolist()):
        px0, px1, py0, py1 = t[4:8]
        ft = {}
        HP.rrdbnet_fp16_storage(m, x[:, :, py0:py1, px0:px1], ft)
        i, j = divmod(ti, tiles_x)
        got = atlas[ay[i]:ay[i] + rowh[i], ax[j]:ax[j] + colw[j]]
        assert rel_l2(got, ft["trunk"]) < 4e-4, (ti, rel_l2(got, ft["trunk"]))
    # gap rows / columns of the atlas are exactly zero (they ARE the per-tile zero padding)
    for j in range(tiles_x - 1):
        assert not atlas[:, ax[j + 1] - 1].any()
    for i in range(len(rowh) - 1):
        assert not atlas[ay[i + 1] - 1].any()


@pytest.mark.parametrize("name,H,W,tile,pad,blend,gain", [
    ("RealESRGAN_x4plus", 96, 120, 1 << 20, 10, "crop", 1.0),
    ("RealESRGAN_x4plus", 96, 120, 64, 16, "gaussian", 10.0),
    ("RealESRGAN_x2plus", 96, 120, 64, 16, "crop", 10.0),
])
def test_image_parity_inrange_amplified(gpu_lib, name, H, W, tile, pad, blend, gain):
    """8-bit frames with weights that neither clamp nor make the body irrelevant: +-1 LSB / PSNR >= 50 dB vs the fp32 oracle."""
    sd = inrange_state_dict(name, seed=0, gain=gain)
    f = synth_frame(H, W, seed=6)
    gpu = _gpu(name, sd, tile, pad, blend)
    out = gpu.process_frame(f)
    gpu.close()
    ref = _oracle(name, sd, tile, pad, blend).process_frame(f)
    sat = float(((ref == 0) | (ref == 255)).mean())
    print(f"\n[inrange] {name} gain {gain}: max {max_lsb(out, ref)} LSB, PSNR {psnr_u8(out, ref):.2f} dB, clamped {sat:.2%}")
    assert sat < 0.01
    assert max_lsb(out, ref) <= LSB_TOL and psnr_u8(out, ref) >= PSNR_TOL and ref.std() > 2.0


def test_feature_check_detects_a_wrong_deep_layer(gpu_lib):
    """The 23-block check is sensitive to gross errors anywhere in the body: the CUDA path is given weights in which one conv
    of block 11 is dropped (zero), or blocks 5 and 17 are swapped, and its features are compared with the fp16-storage model
    of the CORRECT weights. Both are > 3 x the 3e-3 tolerance of test_feature_parity_amplified (CPU calibration: 2.7e-2 and
    3.2e-2), while the default-init 8-bit comparison of round 1 barely moves."""
    name = "RealESRGAN_x4plus"
    sd = inrange_state_dict(name, seed=0, gain=10.0)
    f = synth_frame(72, 88, seed=3)
    dropped = dict(sd)
    dropped["body.11.rdb2.conv3.weight"] = np.zeros_like(sd["body.11.rdb2.conv3.weight"])
    swapped = dict(sd)
    for k in sd:
        if k.startswith("body.5."):
            k2 = k.replace("body.5.", "body.17.", 1)
            swapped[k], swapped[k2] = sd[k2], sd[k]
    for tag, bad in (("one conv dropped", dropped), ("two blocks swapped", swapped)):
        got, _, f16 = _features(name, bad, f, sd_oracle=sd)
        e = rel_l2(got["body"], f16["body"])
        print(f"\n[features] {tag}: rel L2 of the body vs the correct model {e:.2e} (tolerance 3e-3)")
        assert e > 1e-2
