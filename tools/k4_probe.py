"""K4 ablations and knobs, isolated (device-resident hook): which part of the pipeline bounds the fused layer pair.
Usage (GPU box): python tools/k4_probe.py"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from video_restore_b200 import _lib
rng = np.random.default_rng(0)
out = []
for H, W, cin in ((720, 1280, 64), (720, 1280, 128), (848, 1538, 64), (848, 1538, 128)):
    x = (rng.standard_normal((H, W, cin)) * 0.25).astype(np.float32)
    wa = (rng.standard_normal((32, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32)
    wb = (rng.standard_normal((32, cin + 32, 3, 3)) / np.sqrt(9 * (cin + 32))).astype(np.float32)
    b = np.zeros(32, np.float32)
    row = []
    for fl in (0, 4, 4 + 16 + 8, 4 + 16 + 8 + 2, 4 + 16 + 8 + 32, 4 + 16 + 8 + 2 + 32, 2, 32 + 16 + 8):
        _, _, ms = _lib.conv_pair2(x, wa, b, wb, b, iters=20, flags=fl)
        row.append(ms * 1e3)
    out.append("%%dx%%d cin %%3d: full %%6.1f  noMMA %%6.1f  skeleton %%6.1f  skel-noTMA %%6.1f  skel-noTMEM %%6.1f  skel-noTMA-noTMEM %%6.1f | full-noTMA %%6.1f  full-noEpilogueData %%6.1f us" %% (W, H, cin, *row))
print("\n".join(out))
""" % str(ROOT)
for env in ({}, {"VR_UNIT": "2"}, {"VR_UNIT": "1"}):
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True)
    print("==", env or "default (unit 3, lag 2)")
    print(r.stdout.strip() or r.stderr[-2000:])
