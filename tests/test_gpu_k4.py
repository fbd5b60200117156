"""K4 (conv3x3_pair2_sm100.cuh): two consecutive 32-channel dense-block layers in ONE launch (`-m gpu`).

Layer B reads layer A's output through a shared-memory hand-off and A's halo rows / columns are recomputed per CTA, so the
things to pin are: every owned element written exactly (126-pixel strips, band edges, image edges), the hand-off rows / columns
zeroed where the 3x3 conv's zero padding or a tile-atlas gap demands it, rings and hand-off slots carried across work items,
and the whole network agreeing with the K3 path. Reference = torch conv2d in fp32 on fp16-rounded operands (the dense block
behind RRDBNet(...) at reference video_upscaler.py:314-315)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ATOL, RTOL = 2e-3, 2e-3


def _ref(x, wa, ba, wb, bb, gaps_x=(), gaps_y=()):
    import torch
    import torch.nn.functional as F

    h16 = lambda t: t.half().float()
    xt = torch.from_numpy(x).permute(2, 0, 1)[None]
    mask = torch.ones(1, 1, x.shape[0], x.shape[1])
    for g in gaps_x:
        mask[..., :, g] = 0
    for g in gaps_y:
        mask[..., g, :] = 0
    ya = h16(F.leaky_relu(F.conv2d(xt, torch.from_numpy(wa), torch.from_numpy(ba), padding=1), 0.2) * mask)
    yb = F.leaky_relu(F.conv2d(torch.cat((xt, ya), 1), torch.from_numpy(wb), torch.from_numpy(bb), padding=1), 0.2) * mask
    return ya[0].permute(1, 2, 0).numpy(), yb[0].permute(1, 2, 0).numpy()


def _case(H, W, cin, seed=0, gaps_x=(), gaps_y=()):
    from video_restore_b200 import _lib

    rng = np.random.default_rng(seed)
    h16 = lambda a: a.astype(np.float16).astype(np.float32)
    x = h16(rng.standard_normal((H, W, cin)).astype(np.float32))
    for g in gaps_x:
        x[:, g] = 0          # gap positions of the source are zero in the network (every epilogue re-zeroes them)
    for g in gaps_y:
        x[g, :] = 0
    wa = h16((rng.standard_normal((32, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32))
    wb = h16((rng.standard_normal((32, cin + 32, 3, 3)) / np.sqrt(9 * (cin + 32))).astype(np.float32))
    ba = (rng.standard_normal(32) * 0.1).astype(np.float32)
    bb = (rng.standard_normal(32) * 0.1).astype(np.float32)
    ya, yb, _ = _lib.conv_pair2(x, wa, ba, wb, bb, gaps_x=gaps_x, gaps_y=gaps_y)
    ra, rb = _ref(x, wa, ba, wb, bb, gaps_x, gaps_y)
    assert np.isfinite(ya).all() and np.isfinite(yb).all(), "an owned element was never written"
    ea, eb = np.abs(ya - ra), np.abs(yb - rb)
    assert (ea <= ATOL + RTOL * np.abs(ra)).all(), f"layer A: max err {ea.max():.3e} at {np.unravel_index(ea.argmax(), ea.shape)}"
    assert (eb <= ATOL + RTOL * np.abs(rb)).all(), f"layer B: max err {eb.max():.3e} at {np.unravel_index(eb.argmax(), eb.shape)}"
    return ya, yb


@pytest.mark.parametrize("H,W,cin", [(8, 126, 64), (8, 128, 64), (1, 1, 64), (1, 33, 64), (2, 130, 128), (3, 127, 64), (5, 17, 128),
                                     (37, 300, 64), (40, 253, 128), (75, 256, 64), (131, 130, 64), (53, 379, 32), (23, 140, 96)])
def test_pair2_shapes(gpu_lib, H, W, cin):
    _case(H, W, cin)


def test_pair2_frame_size(gpu_lib):
    """720p: 11 strips of 126 pixels x 13 bands on 70 clusters (several work items per cluster at some sizes), plus determinism."""
    a1, b1 = _case(720, 1280, 128, seed=3)
    a2, b2 = _case(720, 1280, 128, seed=3)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2)
    _case(300, 1538, 64, seed=4)   # the 6-tile atlas width: 13 strips


def test_pair2_tile_atlas_gaps(gpu_lib):
    """Gap columns / rows of the tile atlas: zero in both outputs, and -- through the hand-off -- zero as layer B's input."""
    ya, yb = _case(61, 300, 64, seed=5, gaps_x=(100, 201), gaps_y=(30,))
    assert not ya[:, 100].any() and not yb[:, 201].any() and not ya[30].any() and not yb[30].any()
    _case(40, 260, 128, seed=6, gaps_x=(125, 126, 127), gaps_y=(0, 39))   # gaps on strip boundaries and image edges


def test_pair2_several_items_per_cluster(gpu_lib, monkeypatch):
    """Grid capped at 3 clusters: TMEM rings, TMA slots, hand-off slots and every barrier phase carry over between work items."""
    monkeypatch.setenv("VR_MAX_CTAS", "6")
    _case(97, 700, 64, seed=7)
    _case(61, 300, 128, seed=8)
    _case(50, 130, 64, seed=9)


def test_pair2_random_shapes(gpu_lib):
    rng = np.random.default_rng(20261019)
    for i in range(12):
        _case(int(rng.integers(1, 90)), int(rng.integers(1, 420)), int(rng.choice([32, 64, 96, 128])), seed=200 + i)


def test_network_k4_vs_k3(gpu_lib, monkeypatch):
    """Whole network with the layer pairs fused (default) vs separate K3 launches (VR_K4=0): deterministic, and the 8-bit frames
    agree within one level (the two paths sum an output row's taps in different orders)."""
    from util import random_state_dict, synth_frame
    from video_restore_b200.restorer import FrameRestorer

    name = "RealESRGAN_x4plus_anime_6B"
    sd = random_state_dict(name, seed=0)
    f = synth_frame(200, 300, seed=17)

    def run(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        r = FrameRestorer(name, sd, tile=128, tile_pad=16)     # 2x3 tiles -> atlas with gap rows and columns
        out = r.process_frame(f)
        n = r.conv_launch_count
        r.close()
        for k in env:
            monkeypatch.delenv(k)
        return out, n

    k4, n4 = run({"VR_K4": "2"})                  # 2 = always (the default, 1, skips widths where 126-pixel strips cost a strip)
    k3, n3 = run({"VR_K4": "0"})
    assert n4 == n3 - 2 * 18                       # 6 blocks x 3 dense blocks x two pairs
    assert np.array_equal(k4, run({"VR_K4": "2"})[0])
    d = np.abs(k4.astype(np.int32) - k3.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 3e-2
