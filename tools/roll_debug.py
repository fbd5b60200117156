"""Debug helper: error map of one K2 case. python tools/roll_debug.py H W cin cout"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from video_restore_b200 import _lib
from probe_conv import ref_conv, h16
H, W, cin, cout = map(int, sys.argv[1:5])
rng = np.random.default_rng(0)
x = h16(rng.standard_normal((H, W, cin)).astype(np.float32))
w = h16((rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32))
b = rng.standard_normal(cout).astype(np.float32) * 0.1
ref = ref_conv(x, w, b)
for rep in range(3):
    y, _ = _lib.conv3x3(x, w, b, flags=128)
    err = np.abs(y - ref)
    bad = err > 2e-2 + 4e-3 * np.abs(ref)
    rows = np.where(bad.any(axis=(1, 2)))[0]
    print(f"rep {rep}: bad rows {rows.tolist()}")
    for r in rows[:4]:
        cols = np.where(bad[r].any(axis=1))[0]
        chs = np.where(bad[r].any(axis=0))[0]
        print(f"   row {r}: {len(cols)} bad cols [{cols.min()}..{cols.max()}], bad channels {chs.tolist()[:40]}")
        c0 = cols[0]
        print(f"      y[{r},{c0},:4]={y[r, c0, :4]} ref={ref[r, c0, :4]}")
        # is it a missing tap? compare with partial sums of single dy taps
        xp = np.zeros((H + 2, W + 2, cin), np.float32); xp[1:-1, 1:-1] = x
        for drop in range(3):
            part = b.copy()
            for dy in range(3):
                if dy == drop: continue
                for dx in range(3):
                    part = part + xp[r + dy, c0 + dx] @ w[:, :, dy, dx].T
            print(f"      without dy={drop}: {part[:4]}")
