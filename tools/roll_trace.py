"""K2 issuer timeline (CTA 0) for one layer shape: python tools/roll_trace.py cin cout [extra_flags]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from video_restore_b200 import _lib
cin, cout = int(sys.argv[1]), int(sys.argv[2])
fl = 128 + 256 + (int(sys.argv[3]) if len(sys.argv) > 3 else 0)
ms = _lib.conv3x3_bench(720, 1280, cin, cout, rows=0, flags=fl, iters=3)
print(f"{cin}->{cout} flags={fl}: {ms*1e3:.1f} us")
