"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path through the C ABI vs the CPU oracle.

Tolerances (north star): uint8 frames within +-1 LSB and PSNR >= 50 dB for the fp16 network path; bit-exact for tile
indices, CLAHE histogram/LUT and every integer / pinned-float filter.
"""
from pathlib import Path

import numpy as np
import pytest

from util import max_lsb, oracle_model_from_sd, psnr_u8, random_state_dict, synth_frame

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"

LSB_TOL = 1          # uint8 levels
PSNR_TOL = 50.0      # dB
# single conv: fp16-rounded operands, fp32 accumulation, ONE fp16 rounding of the result (half an ulp = 4.9e-4 relative);
# 2e-3 + 2e-3 |ref| is four half-ulps (round 1 allowed 2e-2: ~40)
CONV_ATOL, CONV_RTOL = 2e-3, 2e-3


# ----------------------------------------------------------------------------------------------------------
# K1: single convolution against torch.nn.functional.conv2d (fp32, CPU)
# ----------------------------------------------------------------------------------------------------------
FORCE_TILE, FORCE_ROLL, FORCE_PAIR = 64, 128, 512   # ConvFlags: kernel selection in the conv hook (K1 / K2 / K3)
PLANAR = 1024                                        # ... and chunk-planar source / residual / output tensors


def _conv_case(gpu_lib, H, W, cin, cout, act=0, prelu=False, res=0, rows=0, seed=0, flags=0):
    import torch
    import torch.nn.functional as F

    from video_restore_b200 import _lib

    rng = np.random.default_rng(seed)
    h16 = lambda a: a.astype(np.float16).astype(np.float32)
    x = h16(rng.standard_normal((H, W, cin)).astype(np.float32))
    w = h16((rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32))
    b = (rng.standard_normal(cout) * 0.1).astype(np.float32)
    pr = (rng.random(cout) * 0.5).astype(np.float32) if prelu else None
    r1 = h16(rng.standard_normal((H, W, cout)).astype(np.float32)) if res >= 1 else None
    r2 = h16(rng.standard_normal((H, W, cout)).astype(np.float32)) if res >= 2 else None
    ref = F.conv2d(torch.from_numpy(x).permute(2, 0, 1)[None], torch.from_numpy(w), torch.from_numpy(b), padding=1)
    ref = ref[0].permute(1, 2, 0).numpy()
    if prelu:
        ref = np.where(ref > 0, ref, ref * pr)
    elif act == 1:
        ref = np.where(ref > 0, ref, ref * 0.2)
    if r1 is not None:
        ref = ref * 0.2 + r1
    if r2 is not None:
        ref = ref * 0.2 + r2
    if cout == 48:
        ref = ref.reshape(H, W, 3, 4, 4).transpose(0, 3, 1, 4, 2).reshape(4 * H, 4 * W, 3)
        ref = ref + np.repeat(np.repeat(x[:, :, :3], 4, axis=0), 4, axis=1)
    y, _ = _lib.conv3x3(x, w, b, act=2 if prelu else act, prelu=pr, res1=r1, s1=0.2, res2=r2, s2=0.2, rows=rows,
                        flags=flags)
    err = np.abs(y - ref)
    assert (err <= CONV_ATOL + CONV_RTOL * np.abs(ref)).all(), f"max err {err.max():.3e}"


@pytest.mark.parametrize("H,W,cin,cout", [(8, 128, 32, 32), (37, 300, 64, 32), (16, 256, 96, 32), (16, 256, 160, 32),
                                          (21, 200, 192, 64), (5, 17, 64, 64), (9, 1280, 64, 64), (1, 1, 64, 64),
                                          (12, 140, 3, 64), (12, 140, 12, 64), (12, 140, 64, 3), (12, 140, 64, 48),
                                          (3, 129, 128, 32), (131, 130, 64, 32)])
def test_conv_shapes(gpu_lib, H, W, cin, cout):
    _conv_case(gpu_lib, H, W, cin, cout, flags=FORCE_TILE)      # K1 on every shape
    _conv_case(gpu_lib, H, W, cin, cout)                        # default dispatch (K2 where it is enabled)


# K2 (rolling-row kernel): bands, TMEM ring wrap (16 blocks at 32 channels, 8 at 64), band edges of 1 / 2 / 3 rows, the
# 192 -> 64 layer as two resident halves, partial last strip, several strips x bands > SM count
@pytest.mark.parametrize("H,W,cin,cout", [(8, 128, 32, 32), (75, 128, 64, 32), (75, 200, 64, 64), (37, 300, 64, 32),
                                          (40, 256, 160, 32), (41, 200, 192, 64), (5, 17, 64, 64), (1, 33, 64, 32),
                                          (2, 130, 96, 32), (3, 129, 128, 32), (12, 140, 3, 64), (12, 140, 12, 64),
                                          (131, 130, 64, 32), (300, 1280, 128, 32)])
def test_conv_rolling_shapes(gpu_lib, H, W, cin, cout):
    _conv_case(gpu_lib, H, W, cin, cout, flags=FORCE_ROLL)


@pytest.mark.parametrize("kw", [dict(act=1, cout=32), dict(prelu=True, cout=64), dict(prelu=True, cout=32), dict(res=1, cout=64),
                                dict(res=2, cout=64), dict(res=2, cout=32, act=1)])
def test_conv_rolling_epilogues(gpu_lib, kw):
    kw = dict(kw)
    cout = kw.pop("cout")
    _conv_case(gpu_lib, 23, 140, 192 if cout == 64 and "res" in kw else 64, cout, flags=FORCE_ROLL, **kw)


# K3 (CTA-pair rolling-row kernel, tcgen05 cta_group::2): odd strip counts (empty second strip), mirrored ring several times
# round (period 14 / 6), phantom rows around 1..3-row bands, the resident 192 -> 64 layer, many clusters
@pytest.mark.parametrize("H,W,cin,cout", [(8, 256, 32, 32), (8, 128, 64, 32), (75, 256, 64, 32), (75, 200, 64, 64),
                                          (37, 300, 64, 32), (40, 256, 160, 32), (41, 200, 192, 64), (5, 17, 64, 64),
                                          (1, 33, 64, 32), (2, 130, 96, 32), (3, 129, 128, 32), (12, 140, 3, 64),
                                          (12, 140, 12, 64), (131, 130, 64, 32), (300, 1280, 192, 64)])
def test_conv_pair_shapes(gpu_lib, H, W, cin, cout):
    _conv_case(gpu_lib, H, W, cin, cout, flags=FORCE_PAIR)


@pytest.mark.parametrize("kw", [dict(act=1, cout=32), dict(prelu=True, cout=64), dict(prelu=True, cout=32), dict(res=1, cout=64),
                                dict(res=2, cout=64), dict(res=2, cout=32, act=1)])
def test_conv_pair_epilogues(gpu_lib, kw):
    kw = dict(kw)
    cout = kw.pop("cout")
    _conv_case(gpu_lib, 23, 140, 192 if cout == 64 and "res" in kw else 64, cout, flags=FORCE_PAIR, **kw)


@pytest.mark.parametrize("kernel", [FORCE_TILE, FORCE_ROLL, FORCE_PAIR])
def test_conv_chunk_planar_tensors(gpu_lib, kernel):
    """The network's activation layout: planes of 32 channels ([plane][H][W][32]); source prefix over several planes, output
    into one or two planes, residuals read from planes (the hook converts to / from [H][W][C])."""
    fl = kernel + PLANAR
    _conv_case(gpu_lib, 37, 300, 64, 32, flags=fl)
    _conv_case(gpu_lib, 40, 256, 160, 32, act=1, flags=fl)
    _conv_case(gpu_lib, 41, 200, 192, 64, res=2, flags=fl)
    _conv_case(gpu_lib, 23, 140, 64, 64, prelu=True, res=1, flags=fl)
    _conv_case(gpu_lib, 12, 140, 3, 64, flags=fl)
    if kernel == FORCE_TILE:
        _conv_case(gpu_lib, 12, 140, 64, 3, flags=fl)      # RGB output from a planar source
        _conv_case(gpu_lib, 12, 140, 64, 48, flags=fl)     # pixel shuffle + base


@pytest.mark.parametrize("kernel", [FORCE_ROLL, FORCE_PAIR])
def test_conv_rolling_random_shapes(gpu_lib, kernel):
    """Seeded random layer shapes / epilogues / layouts: band and strip edges, ring phases and mirror positions land differently
    for every (H, W); channel counts are the ones the model zoo produces."""
    rng = np.random.default_rng(20261018)
    for i in range(14):
        cout = int(rng.choice([32, 64]))
        cin = int(rng.choice([3, 12, 32, 64, 96, 128, 160, 192] if cout == 64 else [32, 64, 96, 128, 160]))
        H, W = int(rng.integers(1, 90)), int(rng.integers(1, 420))
        kw = dict(act=int(rng.integers(0, 2)), res=int(rng.integers(0, 3)), seed=100 + i)
        if rng.random() < 0.25:
            kw = dict(prelu=True, res=kw["res"], seed=100 + i)
        fl = kernel + (PLANAR if rng.random() < 0.5 else 0)
        _conv_case(gpu_lib, H, W, cin, cout, flags=fl, **kw)


@pytest.mark.parametrize("flags", [FORCE_ROLL, FORCE_PAIR])
def test_conv_rolling_several_items_per_cta(gpu_lib, monkeypatch, flags):
    """Grid capped at 6 CTAs (3 pairs): every CTA / pair walks several (band, strip) items, so the TMEM ring, the slot ring and
    all barrier phases carry over from item to item (at frame sizes the grids are one item per CTA)."""
    monkeypatch.setenv("VR_MAX_CTAS", "6")
    _conv_case(gpu_lib, 97, 700, 96, 32, flags=flags)
    _conv_case(gpu_lib, 61, 300, 192, 64, res=2, flags=flags)
    _conv_case(gpu_lib, 50, 130, 64, 64, act=1, flags=flags)


def test_conv_rolling_rejects_other_layers(gpu_lib):
    from video_restore_b200._lib import VrError
    with pytest.raises(VrError):
        _conv_case(gpu_lib, 12, 140, 64, 3, flags=FORCE_ROLL)   # RGB output stays on K1
    with pytest.raises(VrError):
        _conv_case(gpu_lib, 12, 140, 64, 48, flags=FORCE_PAIR)  # pixel-shuffle output too


@pytest.mark.parametrize("kw", [dict(act=1), dict(prelu=True), dict(res=1), dict(res=2), dict(act=1, rows=8)])
def test_conv_epilogues(gpu_lib, kw):
    cout = 32 if kw.get("rows") == 8 or kw.get("act") == 1 else 64
    _conv_case(gpu_lib, 12, 140, 64 if cout == 32 else 192, cout, flags=FORCE_TILE, **kw)


# ----------------------------------------------------------------------------------------------------------
# integer / pinned-float stages: bit-exact
# ----------------------------------------------------------------------------------------------------------
def test_tile_grid_bit_exact(gpu_lib):
    from oracle.realesrganer import tile_grid as o_tile_grid
    from video_restore_b200.restorer import tile_grid

    g = np.load(G / "tile_grids.npz")
    for i, c in enumerate(g["cases"].tolist()):
        assert np.array_equal(tile_grid(*c), g[f"grid_{i}"]) and np.array_equal(tile_grid(*c), o_tile_grid(*c))


@pytest.mark.parametrize("shape", [(97, 131), (64, 64), (5, 7), (240, 427), (96, 128), (80, 256), (128, 384), (150, 448), (24, 192)])
def test_filters_bit_exact_vs_oracle(gpu_lib, shape):
    from oracle import filters as OF
    from video_restore_b200 import restorer as R

    f = synth_frame(*shape, seed=1)
    f2 = synth_frame(*shape, seed=1, index=1)
    assert np.array_equal(R.bilateral_filter(f), OF.bilateral_filter(f))
    assert np.array_equal(R.bilateral_filter(f, 7, 40.0, 3.0), OF.bilateral_filter(f, 7, 40.0, 3.0))
    assert np.array_equal(R.unsharp_mask(f, 0.5), OF.unsharp_mask(f, 0.5))
    assert np.array_equal(R.temporal_blend(f2, f), OF.temporal_blend(f2, f))
    assert np.array_equal(R.temporal_blend(f2, f, 0.35, 30.0), OF.temporal_blend(f2, f, 0.35, 30.0))
    for clip, grid in ((2.0, 8), (3.5, 4)):
        if min(shape) < grid:
            continue
        out, hist, lut = R.clahe_bgr(f, clip, grid, return_tables=True)
        oh, ol, _, _ = OF.clahe_tables(np.ascontiguousarray(OF.bgr_to_ycrcb(f)[:, :, 0]), clip, grid)
        assert np.array_equal(hist, oh), "CLAHE histogram (clip + redistribute) must be bit-exact"
        assert np.array_equal(lut, ol), "CLAHE LUT must be bit-exact"
        assert np.array_equal(out, OF.clahe_bgr(f, clip, grid))


def test_filters_vs_cv2_golden(gpu_lib):
    from video_restore_b200 import restorer as R

    g = np.load(G / "filters_cv2.npz")
    d = np.abs(R.bilateral_filter(g["frame"]).astype(np.int32) - g["bilateral"].astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 1e-4   # cv2's own SIMD/scalar paths differ at this level
    s = np.load(G / "filters_spec.npz")
    out, hist, lut = R.clahe_bgr(s["frame"], return_tables=True)
    assert np.array_equal(hist, s["clahe_hist"]) and np.array_equal(lut, s["clahe_lut"])
    assert np.array_equal(out, s["clahe_bgr"])
    assert np.array_equal(R.unsharp_mask(s["frame"], 0.5), s["unsharp"])
    assert np.array_equal(R.temporal_blend(s["frame2"], s["frame"]), s["temporal"])


def test_blend_weights(gpu_lib):
    from oracle.realesrganer import blend_window
    from video_restore_b200.restorer import blend_weights

    for e in (1, 2, 96, 1000, 2304):
        assert np.abs(blend_weights(e) - blend_window(e)).max() <= 2.4e-7   # 2 ulp of expf at <= 1.0


# ----------------------------------------------------------------------------------------------------------
# end to end: FrameRestorer vs OracleRestorer
# ----------------------------------------------------------------------------------------------------------
def _pair(name, tile, pad, blend="crop"):
    from oracle.pipeline import OracleRestorer
    from video_restore_b200.restorer import FrameRestorer

    sd = random_state_dict(name, seed=0)
    return (FrameRestorer(name, sd, tile=tile, tile_pad=pad, blend=blend),
            OracleRestorer(name, tile=tile, tile_pad=pad, blend=blend, model=oracle_model_from_sd(name, sd)))


def _check(out, ref):
    assert out.shape == ref.shape and out.dtype == np.uint8
    assert max_lsb(out, ref) <= LSB_TOL, f"max {max_lsb(out, ref)} LSB"
    assert psnr_u8(out, ref) >= PSNR_TOL, f"PSNR {psnr_u8(out, ref):.2f} dB"
    assert ref.std() > 2.0, "degenerate reference image"


@pytest.mark.parametrize("name,H,W,tile,pad,blend", [
    ("RealESRGAN_x4_v3", 60, 100, 1024, 10, "crop"),
    ("RealESRGAN_x4_v3", 61, 77, 32, 10, "gaussian"),
    ("RealESRGAN_x4plus_anime_6B", 48, 64, 1024, 10, "crop"),
    ("RealESRGAN_x4plus_anime_6B", 150, 37, 64, 16, "crop"),        # thin tiles, W < one conv tile
    ("RealESRGAN_x4plus", 64, 64, 48, 8, "crop"),
    ("RealESRGAN_x4plus", 40, 72, 32, 8, "gaussian"),
    ("RealESRGAN_x2plus", 66, 90, 32, 8, "crop"),
    ("RealESRGAN_x2plus", 65, 91, 32, 8, "gaussian"),               # odd extent: reflect mod-pad
    ("RealESRGAN_x4plus_anime_6B", 40, 150, 16, 4, "crop"),         # 10 tiles per row: two atlas groups (8 + 2 columns)
    ("RealESRGAN_x4_v3", 150, 40, 16, 4, "gaussian"),               # 10 tile rows, blended across the group boundary
    ("RealESRGAN_x2plus", 36, 280, 32, 8, "crop"),                  # 9 tile columns on the pixel-unshuffled grid
])
def test_enhance_parity(gpu_lib, name, H, W, tile, pad, blend):
    gpu, orc = _pair(name, tile, pad, blend)
    f = synth_frame(H, W, seed=3)
    _check(gpu.process_frame(f), orc.process_frame(f))
    gpu.close()


def test_baseline_config1(gpu_lib):
    """BASELINE.json configs[0]: x4plus, 256x256 frame, tile 128 overlap 16, vs the fp32 CPU path."""
    gpu, orc = _pair("RealESRGAN_x4plus", 128, 16)
    f = synth_frame(256, 256, seed=0)
    _check(gpu.process_frame(f), orc.process_frame(f))
    gpu.close()


def test_realesrganer_signature(gpu_lib):
    """The reference's constructor/enhance contract (video_upscaler.py:328-338, :501)."""
    from oracle.realesrganer import RealESRGANer as ORef
    from video_restore_b200.restorer import RealESRGANer

    name = "RealESRGAN_x4plus_anime_6B"
    sd = random_state_dict(name, seed=0)
    up = RealESRGANer(scale=4, model_path=None, model=name, tile=32, tile_pad=10, pre_pad=0, half=True, gpu_id=0,
                      state_dict=sd)
    f = synth_frame(40, 50, seed=2)
    f.setflags(write=False)                     # the reference hands over read-only frames (np.frombuffer, :246)
    out, mode = up.enhance(f, outscale=4)
    ref, _ = ORef(4, oracle_model_from_sd(name, sd), tile=32, tile_pad=10, pre_pad=0).enhance(f, outscale=4)
    assert mode == "RGB"
    _check(out, ref)
    # outscale != scale: upstream's last step, a Lanczos resize of the network-scale result with the same OpenCV call
    import cv2
    half, _ = up.enhance(f, outscale=2)
    assert half.shape == (80, 100, 3)
    assert np.array_equal(half, cv2.resize(out, (100, 80), interpolation=cv2.INTER_LANCZOS4))
    odd, _ = up.enhance(f, outscale=2.5)
    assert odd.shape == (100, 125, 3)
    with pytest.raises(ValueError):
        RealESRGANer(scale=2, model=name, state_dict=sd)
    # upstream's DEFAULT pre_pad = 10 (reflect pad bottom / right, crop afterwards); the reference passes 0 (:334)
    up10 = RealESRGANer(scale=4, model=name, tile=32, tile_pad=10, half=True, gpu_id=0, state_dict=sd)
    assert up10.pre_pad == 10
    out10, _ = up10.enhance(f, outscale=4)
    ref10, _ = ORef(4, oracle_model_from_sd(name, sd), tile=32, tile_pad=10, pre_pad=10).enhance(f, outscale=4)
    _check(out10, ref10)


def test_x2_odd_tile_rejected(gpu_lib):
    from video_restore_b200._lib import VrError
    from video_restore_b200.restorer import FrameRestorer

    r = FrameRestorer("RealESRGAN_x2plus", random_state_dict("RealESRGAN_x2plus", 0), tile=32, tile_pad=7)
    with pytest.raises(VrError, match="even"):
        r.process_frame(synth_frame(70, 70, seed=1))
    r.close()


def test_missing_weights_rejected(gpu_lib):
    from video_restore_b200._lib import VrError
    from video_restore_b200.restorer import FrameRestorer

    sd = random_state_dict("RealESRGAN_x4_v3", 0)
    sd.pop("body.10.bias")
    with pytest.raises(VrError, match="missing tensor"):
        FrameRestorer("RealESRGAN_x4_v3", sd)
    r = FrameRestorer("RealESRGAN_x4_v3", None)
    with pytest.raises(VrError, match="commit_weights"):
        r.process_frame(synth_frame(16, 16, 0))
    r.close()


def test_enhancement_chain_composition(gpu_lib):
    """The full chain on the GPU == oracle filters applied to the GPU's own intermediate frames, bit for bit; the
    network part is within tolerance (previous tests). CLAHE amplifies +-1 LSB network differences on low-contrast
    random-init outputs, so the end-to-end comparison with the oracle chain is a PSNR check only."""
    from oracle import filters as OF
    from oracle.pipeline import FrameOpts as OOpts
    from video_restore_b200.restorer import FrameOpts

    gpu, orc = _pair("RealESRGAN_x4plus_anime_6B", 32, 8, "gaussian")
    frames = [synth_frame(70, 90, seed=5, index=i) for i in range(3)]
    opts = FrameOpts(denoise=True, sharpen=0.5, clahe=True, temporal=True)
    outs = [gpu.process_frame(f, opts) for f in frames]
    ups = [OF.clahe_bgr(OF.unsharp_mask(gpu.process_frame(f, FrameOpts(denoise=True)), 0.5)) for f in frames]
    assert np.array_equal(outs[0], ups[0])
    for t in (1, 2):
        assert np.array_equal(outs[t], OF.temporal_blend(ups[t], ups[t - 1]))
    # temporal state hand-over (what a frame-range shard does at its boundary)
    gpu.temporal_reset()
    gpu.temporal_set_prev(ups[1])
    assert np.array_equal(gpu.process_frame(frames[2], opts), outs[2])
    assert np.array_equal(gpu.temporal_get_prev(*ups[2].shape[:2]), ups[2])
    ref = [orc.process_frame(f, OOpts(denoise=True, sharpen=0.5, clahe=True, temporal=True)) for f in frames]
    assert min(psnr_u8(a, b) for a, b in zip(outs, ref)) > 40.0
    gpu.close()


def test_enhancement_chain_vectorised_fused_temporal(gpu_lib):
    """Frame sizes that take the vectorised kernels (HR width / 8 a multiple of 16): from the second frame on the temporal blend
    is fused into CLAHE's apply pass (one kernel writes the un-blended frame AND the blended one) -- still bit-exact against the
    oracle filters applied to the GPU's own upscaled frames, including the temporal state it leaves behind."""
    from oracle import filters as OF
    from video_restore_b200.restorer import FrameOpts

    gpu, _ = _pair("RealESRGAN_x4_v3", 64, 10)
    frames = [synth_frame(32, 64, seed=6, index=i) for i in range(4)]
    opts = FrameOpts(denoise=True, sharpen=0.5, clahe=True, temporal=True, temporal_tau=40.0)
    ups = [OF.clahe_bgr(OF.unsharp_mask(gpu.process_frame(f, FrameOpts(denoise=True)), 0.5)) for f in frames]
    gpu.temporal_reset()
    n0 = gpu.launch_count
    outs = []
    per_frame = []
    for f in frames:
        outs.append(gpu.process_frame(f, opts))
        per_frame.append(gpu.launch_count - n0)
        n0 = gpu.launch_count
    assert np.array_equal(outs[0], ups[0])
    for t in (1, 2, 3):
        assert np.array_equal(outs[t], OF.temporal_blend(ups[t], ups[t - 1], 0.2, 40.0)), t
        assert not np.array_equal(outs[t], ups[t])            # the blend really acts on this clip
    assert per_frame[1] == per_frame[2] == per_frame[3]      # frame 0: temporal passes through (a copy); then one fused kernel
    assert np.array_equal(gpu.temporal_get_prev(*ups[3].shape[:2]), ups[3])
    # sharpen off: CLAHE + temporal alone take the same fused pass
    gpu.temporal_reset()
    o2 = FrameOpts(clahe=True, temporal=True, temporal_tau=40.0)
    raw = [gpu.process_frame(f) for f in frames[:2]]
    got = [gpu.process_frame(f, o2) for f in frames[:2]]
    e = [OF.clahe_bgr(r) for r in raw]
    assert np.array_equal(got[0], e[0]) and np.array_equal(got[1], OF.temporal_blend(e[1], e[0], 0.2, 40.0))
    gpu.close()


def test_device_path_equals_host_path_and_is_deterministic(gpu_lib):
    import torch

    from video_restore_b200.restorer import FrameOpts

    gpu, _ = _pair("RealESRGAN_x4_v3", 64, 10)
    f = synth_frame(90, 120, seed=8)
    host = gpu.process_frame(f, FrameOpts(denoise=True, sharpen=0.2))
    d_in = torch.from_numpy(f).cuda()
    d_out = torch.empty((360, 480, 3), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        gpu.process_frame_device(d_in.data_ptr(), 90, 120, d_out.data_ptr(), FrameOpts(denoise=True, sharpen=0.2))
        assert np.array_equal(d_out.cpu().numpy(), host)
    assert gpu.launch_count > 0 and gpu.last_timing()[0] > 0
    gpu.close()


def test_full_size_480p_srvgg_crop_property(gpu_lib):
    """BASELINE configs[1] size (854x480, x4_v3, single tile): the oracle is run on a crop whose margin exceeds the
    network's receptive field (34 convs -> 34 px), so the interior must agree with the full-frame GPU result."""
    gpu, orc = _pair("RealESRGAN_x4_v3", 1024, 10)
    f = synth_frame(480, 854, seed=12)
    out = gpu.process_frame(f)
    assert out.shape == (1920, 3416, 3)
    y0, x0, m, sz = 200, 400, 36, 64
    crop = np.ascontiguousarray(f[y0 - m:y0 + sz + m, x0 - m:x0 + sz + m])
    ref = orc.process_frame(crop)[m * 4:(m + sz) * 4, m * 4:(m + sz) * 4]
    _check(out[y0 * 4:(y0 + sz) * 4, x0 * 4:(x0 + sz) * 4], ref)
    # frame corner: borders are zero padded identically when the crop shares the corner
    crop = np.ascontiguousarray(f[:sz + m, :sz + m])
    _check(out[:sz * 4, :sz * 4], orc.process_frame(crop)[:sz * 4, :sz * 4])
    gpu.close()


def test_full_size_720p_x4plus_runs_and_is_tile_consistent(gpu_lib):
    """BASELINE configs[3] size, properties only (the oracle comparison over the whole frame at this size is
    tests/test_gpu_fullsize.py): determinism, and crop-merge tiling vs single tile differing by <= 1 LSB on < 4 % of pixels
    (tile borders only matter through zero padding 10+ px away; K2 / K3 sum an output row's taps in an
    order that depends on the row's parity inside its band (and, in K3, on its ring position), so identical tile interiors
    can round differently in fp16: 0.5 % with K1 only, 1.2 % with K2 on the 32-channel layers, 2.4 % with K3 everywhere)."""
    from video_restore_b200.restorer import FrameRestorer

    sd = random_state_dict("RealESRGAN_x4plus", seed=0)
    f = synth_frame(720, 1280, seed=13)
    one = FrameRestorer("RealESRGAN_x4plus", sd, tile=1536, tile_pad=10)
    a = one.process_frame(f)
    assert np.array_equal(a, one.process_frame(f))
    one.close()
    tiled = FrameRestorer("RealESRGAN_x4plus", sd, tile=512, tile_pad=64)
    b = tiled.process_frame(f)
    tiled.close()
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    assert a.shape == (2880, 5120, 3) and d.max() <= 1 and (d > 0).mean() < 4e-2
    assert a.std() > 2.0


@pytest.mark.parametrize("H,W,tile,pad", [(90, 132, 32, 8), (61, 77, 32, 10), (150, 40, 16, 4), (64, 200, 48, 10), (37, 53, 64, 6)])
@pytest.mark.parametrize("blend", ["gaussian", "crop"])
def test_merge_fast_kernels_are_bit_identical(gpu_lib, monkeypatch, H, W, tile, pad, blend):
    """The aligned tile-merge kernels (Gaussian blend: 4 pixels per thread, per-tile work hoisted, multiply-high division;
    crop: 4 pixels per thread) against the general per-pixel kernels (VR_BLEND_FAST=0) on the same network output: same
    tiles in the same order, same roundings."""
    from video_restore_b200.restorer import FrameRestorer

    name = "RealESRGAN_x4_v3"
    sd = random_state_dict(name, seed=3)
    f = synth_frame(H, W, seed=5)
    outs = []
    for fast in ("1", "0"):
        monkeypatch.setenv("VR_BLEND_FAST", fast)
        r = FrameRestorer(name, sd, tile=tile, tile_pad=pad, blend=blend)
        outs.append(r.process_frame(f))
        r.close()
    monkeypatch.delenv("VR_BLEND_FAST")
    assert outs[0].shape == (4 * H, 4 * W, 3)
    assert np.array_equal(outs[0], outs[1])


def test_execution_variants_are_bit_identical(gpu_lib, monkeypatch):
    """Scheduling variants of the conv stack (read from the environment at handle creation) must not change a single
    bit: multi-layer persistent launch (tile-row dependency counters), no PDL, streamed instead of resident weights,
    and row-band scheduling."""
    from video_restore_b200.restorer import FrameRestorer

    name = "RealESRGAN_x4plus_anime_6B"
    sd = random_state_dict(name, seed=0)
    f = synth_frame(200, 300, seed=17)

    def run(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        r = FrameRestorer(name, sd, tile=128, tile_pad=16)     # 2x3 tiles -> atlas with gap rows and columns
        out = r.process_frame(f)
        n = r.launch_count
        r.close()
        for k in env:
            monkeypatch.delenv(k)
        return out, n

    # K1 only (VR_ROLL=0): every scheduling variant is bit-identical
    base, n_base = run({"VR_ROLL": "0"})
    multi, n_multi = run({"VR_ROLL": "0", "VR_MULTI": "1"})
    assert np.array_equal(base, multi) and n_multi < n_base
    assert np.array_equal(base, run({"VR_ROLL": "0", "VR_PDL": "0"})[0])
    assert np.array_equal(base, run({"VR_ROLL": "0", "VR_WRES": "0"})[0])
    # the four phases of each folded upsample conv as four launches instead of one launch walking all phases per input tile
    ph4, n_ph4 = run({"VR_ROLL": "0", "VR_PHASES1": "0"})
    assert np.array_equal(base, ph4) and n_ph4 == n_base + 6
    # K2 / K3 on their layer classes: deterministic and independent of PDL; against K1 the fp32 summation order differs (bias is the
    # accumulator's initial value, taps are summed row by row), so the 8-bit frames agree within one level
    # interleaved instead of chunk-planar activation tensors: same arithmetic in the same order
    assert np.array_equal(base, run({"VR_ROLL": "0", "VR_PLANAR": "0"})[0])
    # K3 everywhere (VR_K4=0: the dense block's layer pairs as separate launches; K4 itself is tests/test_gpu_k4.py)
    k3 = run({"VR_K4": "0"})[0]
    assert np.array_equal(k3, run({"VR_K4": "0", "VR_PLANAR": "0"})[0])
    # K3 epilogue variants: staged through shared memory instead of direct 256-bit stores, ring position handed back after the
    # stores / before them / right after the TMEM loads -- the arithmetic and its order are the same
    for env in ({"VR_EPI_DIRECT": "0"}, {"VR_EARLY64": "0"}, {"VR_EARLY64": "1"}, {"VR_EPI_DIRECT": "0", "VR_EARLY64": "0"}):
        assert np.array_equal(k3, run({"VR_K4": "0", **env})[0]), env
    # the default path (K4 pairs + K3): the same switches only touch its K3 layers -- still bit-identical to itself
    k4 = run({"VR_K4": "2"})[0]
    for env in ({"VR_EARLY64": "0"}, {"VR_EARLY64": "1"}, {"VR_PDL": "0"}, {"VR_PHASES1": "0"}):
        assert np.array_equal(k4, run({"VR_K4": "2", **env})[0]), env
    d = np.abs(k4.astype(np.int32) - k3.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 3e-2
    # smaller tile-atlas groups (VR_ATLAS_TILES bounds activation memory like --tile-size does in the reference): the grid is run
    # as several atlases. K1 is position-independent: bit-identical; the rolling-row kernels within one level
    assert np.array_equal(base, run({"VR_ROLL": "0", "VR_ATLAS_TILES": "1"})[0])
    assert np.array_equal(base, run({"VR_ROLL": "0", "VR_ATLAS_TILES": "2"})[0])
    # automatic group sizing: a memory budget too small for the whole atlas splits it into groups by itself (one 160 x 160 tile
    # of this frame needs ~155 MB of activations), a generous one changes nothing
    small, n_small = run({"VR_ROLL": "0", "VR_MEM_BUDGET_MB": "200"})
    assert np.array_equal(base, small) and n_small > 3 * n_base
    assert run({"VR_ROLL": "0", "VR_MEM_BUDGET_MB": "100000"})[1] == n_base
    g2 = run({"VR_ATLAS_TILES": "2"})[0]
    d = np.abs(g2.astype(np.int32) - k4.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 3e-2
    # a different issuer hand-over granularity moves the points where the two issuing warps alternate; MMAs of different
    # warps into one accumulator are applied in a different order then (measured: not bit-identical), within tolerance
    u1 = run({"VR_K4": "0", "VR_UNIT": "1"})[0]
    assert np.array_equal(u1, run({"VR_K4": "0", "VR_UNIT": "1"})[0])
    d = np.abs(u1.astype(np.int32) - k3.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 2e-2
    for mask in ("1", "7", "24", "31"):
        roll = run({"VR_ROLL": mask})[0]
        assert np.array_equal(roll, run({"VR_ROLL": mask})[0])
        assert np.array_equal(roll, run({"VR_ROLL": mask, "VR_PDL": "0"})[0])
        d = np.abs(roll.astype(np.int32) - base.astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 2e-2


def test_process_stream_equals_per_frame_calls(gpu_lib):
    """Pipelined host path (vr_submit / vr_wait) == one synchronous vr_restore per frame, temporal state included."""
    from video_restore_b200.restorer import FrameOpts

    gpu, _ = _pair("RealESRGAN_x4_v3", 64, 10)
    frames = [synth_frame(72, 100, seed=31, index=i) for i in range(7)]
    opts = FrameOpts(denoise=True, sharpen=0.3, clahe=True, temporal=True)
    ref = [gpu.process_frame(f, opts) for f in frames]
    gpu.temporal_reset()
    got = [o.copy() for o in gpu.process_stream(iter(frames), opts)]
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    gpu.close()


def test_pth_checkpoint_loading(gpu_lib, tmp_path):
    """a .pth with `params_ema` (upstream key names) loads through the RealESRGANer-shaped constructor."""
    import torch

    from video_restore_b200.restorer import RealESRGANer

    name = "RealESRGAN_x4_v3"
    sd = random_state_dict(name, seed=0)
    path = tmp_path / "realesr-general-x4v3.pth"
    torch.save({"params_ema": {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}}, path)
    f = synth_frame(40, 56, seed=4)
    a, _ = RealESRGANer(scale=4, model_path=str(path), model=name, tile=32, tile_pad=10, pre_pad=0, half=True).enhance(f, 4)
    b, _ = RealESRGANer(scale=4, model=name, tile=32, tile_pad=10, pre_pad=0, half=True, state_dict=sd).enhance(f, 4)
    assert np.array_equal(a, b)


def test_cli_synthetic_run(gpu_lib, capsys):
    from video_restore_b200.cli import main

    assert main(["in.mp4", "out.mp4", "--model", "RealESRGAN_x4_v3", "--quality", "fast", "--enhanced",
                 "--synthetic", "2"]) == 0
    assert "processed 2 frames" in capsys.readouterr().out


# ----------------------------------------------------------------------------------------------------------
# size-independent properties at BASELINE full sizes (the oracle is too slow there)
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", [FORCE_TILE, FORCE_PAIR])
@pytest.mark.parametrize("cin,cout", [(160, 32), (192, 64)])
def test_conv_720p_scaling_by_powers_of_two_is_exact(gpu_lib, kernel, cin, cout):
    """Linearity at the headline size (720 x 1280, the dense block's widest layers): multiplying the input by 2 and the
    bias by 2 is exact in fp16 / fp32, so every output must double bit for bit -- whatever the summation order."""
    from video_restore_b200 import _lib

    rng = np.random.default_rng(21)
    H, W = 720, 1280
    x = (rng.standard_normal((H, W, cin)) * 0.25).astype(np.float16).astype(np.float32)
    w = (rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float16).astype(np.float32)
    b = (rng.standard_normal(cout) * 0.1).astype(np.float32)
    y1, _ = _lib.conv3x3(x, w, b, flags=kernel)
    y2, _ = _lib.conv3x3(x * 2.0, w, b * 2.0, flags=kernel)
    assert np.isfinite(y1).all() and y1.std() > 0.1
    normal = np.abs(y1) >= 2.0 ** -13  # below that the fp16 result is subnormal: fixed spacing, doubling is not exact
    assert normal.mean() > 0.999
    assert np.array_equal(y2[normal], y1[normal] * 2.0)
    assert np.abs(y2[~normal] - 2.0 * y1[~normal]).max(initial=0.0) <= 2.0 ** -23
    # and a second run of the same launch is bit-identical (no race between the 140 CTAs / 70 clusters)
    y1b, _ = _lib.conv3x3(x, w, b, flags=kernel)
    assert np.array_equal(y1, y1b)


def test_filters_2880p_properties(gpu_lib):
    """Enhancement kernels on a 5120 x 2880 frame (configs[3] output size): identities and integer invariants."""
    from video_restore_b200 import restorer as R

    small = synth_frame(720, 1280, seed=31)
    f = np.ascontiguousarray(np.repeat(np.repeat(small, 4, axis=0), 4, axis=1))
    f[::7, ::5] = 255 - f[::7, ::5]  # break the 4 x 4 blocks
    assert f.shape == (2880, 5120, 3)
    # unsharp with amount 0 is the identity ((1 + 0) x - 0 blur, rounded)
    assert np.array_equal(R.unsharp_mask(f, 0.0), f)
    # a flat frame is a fixed point of the blur, hence of the unsharp mask, and of the bilateral filter
    flat = np.full((2880, 5120, 3), 77, np.uint8)
    assert np.array_equal(R.unsharp_mask(flat, 0.7), flat)
    assert np.array_equal(R.bilateral_filter(flat[:720, :1280]), flat[:720, :1280])
    # temporal: prev == cur passes through; a frame further than tau away everywhere is not blended
    assert np.array_equal(R.temporal_blend(f, f), f)
    far = (f.astype(np.int32) + 128) % 256
    assert np.array_equal(R.temporal_blend(f, far.astype(np.uint8), 0.2, 12.0), f)
    # CLAHE: every tile histogram sums to the tile area after clipping + redistribution; LUTs are monotone
    out, hist, lut = R.clahe_bgr(f, 2.0, 8, return_tables=True)
    assert hist.shape == (64, 256) and (hist.sum(axis=1) == (2880 // 8) * (5120 // 8)).all()
    assert (np.diff(lut.astype(np.int32), axis=1) >= 0).all()
    assert out.shape == f.shape and out.std() > 0


def test_tile_grid_full_sizes(gpu_lib):
    """Tile tables (integer, bit-exact) at every BASELINE frame size / preset, against the oracle's tile loop."""
    from oracle.realesrganer import tile_grid as o_tile_grid
    from video_restore_b200 import restorer as R

    for H, W, tile, pad, s in [(720, 1280, 512, 64, 4), (720, 1280, 1536, 10, 4), (1080, 1920, 512, 32, 2),
                               (1080, 1920, 1024, 10, 4), (1080, 1920, 512, 32, 4), (480, 854, 1024, 16, 4),
                               (256, 256, 128, 16, 4), (2160, 3840, 512, 32, 4), (2160, 3840, 256, 10, 4),
                               (1080, 1920, 128, 16, 4)]:
        assert np.array_equal(R.tile_grid(H, W, tile, pad, s), o_tile_grid(H, W, tile, pad, s)), (H, W, tile)
