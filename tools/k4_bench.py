"""K4 vs two K3 launches on the dense block's layer pairs, isolated (device-resident hooks, burst clocks), 720p and the 6-tile
atlas size. Usage (GPU box): python tools/k4_bench.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import os
os.environ["VR_BENCH_PLANAR"] = "1"   # the K3 comparison on chunk-planar tensors, as in the network
from video_restore_b200 import _lib  # noqa: E402

rng = np.random.default_rng(0)
for H, W in ((720, 1280), (848, 1538)):
    for cin in (64, 128):
        x = (rng.standard_normal((H, W, cin)) * 0.25).astype(np.float32)
        wa = (rng.standard_normal((32, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32)
        wb = (rng.standard_normal((32, cin + 32, 3, 3)) / np.sqrt(9 * (cin + 32))).astype(np.float32)
        b = np.zeros(32, np.float32)
        _, _, ms = _lib.conv_pair2(x, wa, b, wb, b, iters=30)
        k3a = _lib.conv3x3_bench(H, W, cin, 32, flags=512, iters=30)
        k3b = _lib.conv3x3_bench(H, W, cin + 32, 32, flags=512, iters=30)
        fl = 2 * 9 * (cin * 32 + (cin + 32) * 32) * H * W
        print(f"{W}x{H} conv {cin}->32 + {cin + 32}->32: K4 {ms * 1e3:7.1f} us ({fl / ms / 1e9:6.0f} TFLOP/s)   "
              f"K3 {k3a * 1e3:6.1f} + {k3b * 1e3:6.1f} = {(k3a + k3b) * 1e3:7.1f} us ({fl / (k3a + k3b) / 1e9:6.0f} TFLOP/s)")
