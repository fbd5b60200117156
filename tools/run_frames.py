"""Run N frames of a bench workload through the device path (for ncu / compute-sanitizer captures)."""
import argparse, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from bench import WORKLOADS
from video_restore_b200.restorer import FrameOpts, FrameRestorer
from video_restore_b200.synth import random_state_dict, synth_frame

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4_x4plus_720p_qmax_plain")
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--height", type=int, default=0)
ap.add_argument("--width", type=int, default=0)
a = ap.parse_args()
wl = dict(WORKLOADS[a.workload])
if a.height: wl["H"] = a.height
if a.width: wl["W"] = a.width
s = 2 if "x2" in wl["model"] else 4
r = FrameRestorer(wl["model"], random_state_dict(wl["model"], 0), tile=wl["tile"], tile_pad=wl["pad"], blend=wl["blend"])
d_in = torch.from_numpy(synth_frame(wl["H"], wl["W"], seed=11)).cuda()
d_out = torch.empty((wl["H"] * s, wl["W"] * s, 3), dtype=torch.uint8, device="cuda")
for i in range(a.frames):
    r.process_frame_device(d_in.data_ptr(), wl["H"], wl["W"], d_out.data_ptr(), FrameOpts(**wl["opts"]))
    print("frame", i, "timing", r.last_timing(), "launches", r.launch_count, flush=True)
import hashlib
print("sha1", hashlib.sha1(d_out.cpu().numpy().tobytes()).hexdigest())
