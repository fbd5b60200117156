import sys
sys.path.insert(0, '.')
from video_restore_b200 import _lib
for cin, cout in [(64, 32), (160, 32), (192, 64), (64, 64)]:
    r = {}
    for H in (360, 720, 1440):
        ms = min(_lib.conv3x3_bench(H, 1280, cin, cout, rows=0, flags=512, iters=30) for _ in range(2))
        r[H] = ms * 1e3
    print(f"[ovh] {cin}->{cout}: H=360 {r[360]:.1f} us  H=720 {r[720]:.1f} us  H=1440 {r[1440]:.1f} us  per-launch fixed ~ {2*r[720]-r[1440]:.1f} us (720 vs 1440), {2*r[360]-r[720]:.1f} us (360 vs 720)", flush=True)
