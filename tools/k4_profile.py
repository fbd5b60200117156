"""Where K4's warps wait: cycle counters of cluster 0's leader CTA (issuer warps, producer, two epilogue warps).
Usage (GPU box): python tools/k4_profile.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from video_restore_b200 import _lib  # noqa: E402

rng = np.random.default_rng(0)
for H, W, cin in ((720, 1280, 64), (720, 1280, 128), (848, 1538, 64)):
    x = (rng.standard_normal((H, W, cin)) * 0.25).astype(np.float32)
    wa = (rng.standard_normal((32, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32)
    wb = (rng.standard_normal((32, cin + 32, 3, 3)) / np.sqrt(9 * (cin + 32))).astype(np.float32)
    b = np.zeros(32, np.float32)
    for fl, tag in ((0, "full"), (4, "skip-MMA")):
        _, _, ms = _lib.conv_pair2(x, wa, b, wb, b, iters=5, flags=fl)
        p = _lib.pair2_profile()
        print(f"== {W}x{H} cin {cin} {tag}: {ms * 1e3:.1f} us/launch")
        for k in sorted(p):
            print(f"   {k:24s} {p[k]:10d}")
