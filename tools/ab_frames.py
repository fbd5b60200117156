"""In-process A/B timing of two handle configurations (env vars read at vr_create), frames interleaved ABAB...
usage: python tools/ab_frames.py WORKLOAD "VR_PDL=0" "VR_PDL=1" [rounds]"""
import os, sys, statistics
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from bench import WORKLOADS
from video_restore_b200.restorer import FrameOpts, FrameRestorer
from video_restore_b200.synth import random_state_dict, synth_frame

wl = WORKLOADS[sys.argv[1]]
envs = [dict(kv.split("=") for kv in e.split(",") if kv) for e in sys.argv[2:4]]
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 8
s = 2 if "x2" in wl["model"] else 4
sd = random_state_dict(wl["model"], 0)
rs = []
for env in envs:
    for k, v in env.items():
        os.environ[k] = v
    rs.append(FrameRestorer(wl["model"], sd, tile=wl["tile"], tile_pad=wl["pad"], blend=wl["blend"]))
    for k in env:
        os.environ.pop(k, None)
d_in = torch.from_numpy(synth_frame(wl["H"], wl["W"], seed=11)).cuda()
d_out = torch.empty((wl["H"] * s, wl["W"] * s, 3), dtype=torch.uint8, device="cuda")
opts = FrameOpts(**wl["opts"])
times = [[], []]
for i in range(rounds + 2):
    for j, r in enumerate(rs):
        r.process_frame_device(d_in.data_ptr(), wl["H"], wl["W"], d_out.data_ptr(), opts)
        if i >= 2:
            times[j].append(r.last_timing()[0])
for env, t in zip(envs, times):
    print(f"{env}: median {statistics.median(t):.2f} ms  min {min(t):.2f}  max {max(t):.2f}  (n={len(t)})", flush=True)
