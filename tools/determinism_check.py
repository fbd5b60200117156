"""Run one frame N times on one restorer (and once on a fresh one) and compare sha1 of the outputs: the conv kernels' two issuing
warps feed one tensor pipe, so bit-stability across runs is worth checking under load.  python tools/determinism_check.py [N]"""
import hashlib, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from video_restore_b200.restorer import FrameOpts, FrameRestorer
from video_restore_b200.synth import random_state_dict, synth_frame

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for name, tile, pad, blend, H, W in (("RealESRGAN_x4plus", 1536, 10, "crop", 720, 1280), ("RealESRGAN_x4plus", 512, 64, "gaussian", 720, 1280),
                                     ("RealESRGAN_x4_v3", 1024, 10, "crop", 480, 854)):
    sd = random_state_dict(name, 0)
    hashes = set()
    for rep in range(2):
        r = FrameRestorer(name, sd, tile=tile, tile_pad=pad, blend=blend)
        d_in = torch.from_numpy(synth_frame(H, W, seed=11)).cuda()
        d_out = torch.empty((H * 4, W * 4, 3), dtype=torch.uint8, device="cuda")
        for i in range(n if rep == 0 else 3):
            d_out.zero_()
            r.process_frame_device(d_in.data_ptr(), H, W, d_out.data_ptr(), FrameOpts())
            hashes.add(hashlib.sha1(d_out.cpu().numpy().tobytes()).hexdigest())
        r.close()
    print(f"[determinism] {name} tile {tile}/{pad} {blend} {W}x{H}: {n + 3} runs on two handles -> {len(hashes)} distinct output(s)", flush=True)
