"""Per-launch fixed cost of K4 / K3: t(H) against t(2H) (same strips, bands twice as long): fixed = 2 t(H) - t(2H).
Usage (GPU box): python tools/k4_fixed.py"""
import os
import sys
from pathlib import Path

import numpy as np

os.environ["VR_BENCH_PLANAR"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from video_restore_b200 import _lib  # noqa: E402

rng = np.random.default_rng(0)
W = 1538
for cin in (64, 128):
    t = {}
    for H in (424, 848, 1696):
        x = (rng.standard_normal((H, W, cin)) * 0.25).astype(np.float32)
        wa = (rng.standard_normal((32, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32)
        wb = (rng.standard_normal((32, cin + 32, 3, 3)) / np.sqrt(9 * (cin + 32))).astype(np.float32)
        b = np.zeros(32, np.float32)
        _, _, ms = _lib.conv_pair2(x, wa, b, wb, b, iters=30)
        k3 = _lib.conv3x3_bench(H, W, cin, 32, flags=512, iters=30) + _lib.conv3x3_bench(H, W, cin + 32, 32, flags=512, iters=30)
        t[H] = (ms * 1e3, k3 * 1e3)
        print(f"cin {cin} {W}x{H}: K4 {ms * 1e3:7.1f} us   2 x K3 {k3 * 1e3:7.1f} us")
    for a, b2 in ((424, 848), (848, 1696)):
        print(f"   fixed cost from {a}/{b2} rows: K4 {2 * t[a][0] - t[b2][0]:6.1f} us per launch, K3 pair {2 * t[a][1] - t[b2][1]:6.1f} us (two launches)")
