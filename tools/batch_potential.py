"""What would packing two frames into one tile atlas buy? Time a frame of double height (= the atlas of two frames minus
its gap row) against two frames of the normal height: python tools/batch_potential.py WORKLOAD [frames]"""
import statistics, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from bench import WORKLOADS
from video_restore_b200.restorer import FrameOpts, FrameRestorer
from video_restore_b200.synth import random_state_dict, synth_frame
wl = WORKLOADS[sys.argv[1]]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
s = 2 if "x2" in wl["model"] else 4
sd = random_state_dict(wl["model"], 0)
for mult in (1, 2, 1, 2):
    H, W = wl["H"] * mult, wl["W"]
    tile = wl["tile"] * (mult if wl["tile"] >= wl["H"] else 1)
    r = FrameRestorer(wl["model"], sd, tile=max(tile, 1), tile_pad=wl["pad"], blend=wl["blend"])
    d_in = torch.from_numpy(synth_frame(H, W, seed=11)).cuda()
    d_out = torch.empty((H * s, W * s, 3), dtype=torch.uint8, device="cuda")
    t = []
    for i in range(n // mult):
        r.process_frame_device(d_in.data_ptr(), H, W, d_out.data_ptr(), FrameOpts())
        t.append(r.last_timing()[0])
    r.close()
    med = statistics.median(t[len(t) // 2:])
    print(f"[batch] {sys.argv[1]} height x{mult}: {med:.3f} ms per launch set = {med / mult:.3f} ms per {wl['H']}-row frame (min {min(t) / mult:.3f})", flush=True)
