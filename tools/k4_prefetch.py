"""K4: L2 prefetch distance of layer A rows (row pairs), isolated timing. Usage (GPU box): python tools/k4_lag.py"""
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
row = []
for H, W, cin in ((720, 1280, 64), (720, 1280, 128), (848, 1538, 64), (848, 1538, 128)):
    x = (rng.standard_normal((H, W, cin)) * 0.25).astype(np.float32)
    wa = (rng.standard_normal((32, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32)
    wb = (rng.standard_normal((32, cin + 32, 3, 3)) / np.sqrt(9 * (cin + 32))).astype(np.float32)
    b = np.zeros(32, np.float32)
    _, _, ms = _lib.conv_pair2(x, wa, b, wb, b, iters=30)
    row.append("%%dx%%d/%%d %%6.1f" %% (W, H, cin, ms * 1e3))
print("   ".join(row))
""" % str(ROOT)
for lag in ("0", "2", "3", "5", "8"):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, VR_K4_PREFETCH=lag), capture_output=True, text=True)
    print("prefetch", lag, ":", r.stdout.strip() or r.stderr[-1500:])
