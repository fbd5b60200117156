"""Sustained frame time of one workload under the current environment: python tools/env_frames.py WORKLOAD [frames]"""
import sys, statistics
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from bench import WORKLOADS
from video_restore_b200.restorer import FrameOpts, FrameRestorer
from video_restore_b200.synth import random_state_dict, synth_frame
wl = WORKLOADS[sys.argv[1]]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
s = 2 if "x2" in wl["model"] else 4
r = FrameRestorer(wl["model"], random_state_dict(wl["model"], 0), tile=wl["tile"], tile_pad=wl["pad"], blend=wl["blend"])
d_in = torch.from_numpy(synth_frame(wl["H"], wl["W"], seed=11)).cuda()
d_out = torch.empty((wl["H"] * s, wl["W"] * s, 3), dtype=torch.uint8, device="cuda")
opts = FrameOpts(**wl["opts"])
t = []
for i in range(n):
    r.process_frame_device(d_in.data_ptr(), wl["H"], wl["W"], d_out.data_ptr(), opts)
    t.append(r.last_timing()[0])
print(f"{sys.argv[1]}: first {t[0]:.2f} ms, median of last {n//2}: {statistics.median(t[n//2:]):.2f} ms, min {min(t):.2f}")
