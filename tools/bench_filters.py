"""HBM-bound kernels at BASELINE sizes (run with gpurun): ms and achieved GB/s on algorithmic bytes."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from video_restore_b200 import _lib
root = Path(__file__).resolve().parents[1]
peak = json.loads((root / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (root / "MEASURED_PEAKS.json").exists() else 6650.0
cases = [("bilateral", 720, 1280), ("pre", 720, 1280), ("unsharp", 2880, 5120), ("clahe", 2880, 5120),
         ("temporal", 2880, 5120), ("post_crop", 2880, 5120), ("post_blend", 2880, 5120), ("upsample2x", 1440, 2560)]
only = sys.argv[1:] 
for kind, H, W in cases:
    if only and kind not in only:
        continue
    try:
        ms, gbs = _lib.filter_bench(kind, H, W, iters=20)
        print(f"[filter] {kind:11s} {W}x{H}: {ms*1e3:8.1f} us  {gbs:7.1f} GB/s algorithmic  = {gbs/peak:.3f} of measured copy peak {peak:.0f}", flush=True)
    except Exception as e:
        print(f"[filter] {kind}: ERROR {e}", flush=True)
        break
