"""SECONDARY baseline (never on the product path, not used by bench.py): what a user of the reference gets on this box TODAY --
the reference's own arithmetic (RRDBNet / SRVGGNetCompact in PyTorch, `half=True`, cuDNN with `cudnn.benchmark = True` as
video_upscaler.py:108 sets it, RealESRGANer's tile loop with crop-merge) on the same B200, same weights, same frames.
The networks are the oracle's restatement of basicsr / realesrgan (those packages are not installed here); everything runs in
torch on the GPU, tiles one after the other as upstream's tile_process does, fp32 result copied to the host per frame as
upstream's enhance() does (`.float().cpu()`).

    python tools/torch_gpu_baseline.py [--workload NAME] [--frames N]      (GPU box)
Prints frames/s next to the same workload through libvrb200 (FrameRestorer.process_frame, host frame in / host frame out)."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from bench import WORKLOADS  # noqa: E402
from oracle.realesrganer import tile_grid  # noqa: E402
from util import oracle_model_from_sd  # noqa: E402
from video_restore_b200.restorer import FrameOpts, FrameRestorer  # noqa: E402
from video_restore_b200.synth import random_state_dict, synth_frame  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4_x4plus_720p_qmax_enhanced")
ap.add_argument("--frames", type=int, default=6)
a = ap.parse_args()
wl = WORKLOADS[a.workload]
torch.backends.cudnn.benchmark = True            # video_upscaler.py:108
torch.backends.cuda.matmul.allow_tf32 = True     # :109-110 (irrelevant for the fp16 path)
sd = random_state_dict(wl["model"], 0)
model = oracle_model_from_sd(wl["model"], sd).half().cuda().eval()
s = model.scale if hasattr(model, "scale") else 4
H, W, tile, pad = wl["H"], wl["W"], wl["tile"], wl["pad"]
frames = [synth_frame(H, W, seed=11, index=i) for i in range(2)]


@torch.no_grad()
def enhance(frame):
    img = torch.from_numpy(np.ascontiguousarray(frame[:, :, ::-1].astype(np.float32) / 255.0)).permute(2, 0, 1)[None].cuda().half()
    if s == 2 and (H % 2 or W % 2):
        img = torch.nn.functional.pad(img, (0, W % 2, 0, H % 2), "reflect")
    out = img.new_zeros((1, 3, img.shape[2] * s, img.shape[3] * s))
    for (ix0, ix1, iy0, iy1, px0, px1, py0, py1, ox0, ox1, oy0, oy1) in tile_grid(img.shape[2], img.shape[3], tile, pad, s).tolist():
        t = model(img[:, :, py0:py1, px0:px1])
        out[:, :, iy0 * s:iy1 * s, ix0 * s:ix1 * s] = t[:, :, oy0:oy1, ox0:ox1]
    o = out[0, :, :H * s, :W * s].float().cpu().clamp_(0, 1).numpy()
    return (np.transpose(o[[2, 1, 0]], (1, 2, 0)) * 255.0).round().astype(np.uint8)


for f in frames:
    ref = enhance(f)      # warm-up: cuDNN autotuning per tile shape
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(a.frames):
    ref = enhance(frames[i % 2])
torch.cuda.synchronize()
torch_fps = a.frames / (time.perf_counter() - t0)

r = FrameRestorer(wl["model"], sd, tile=tile, tile_pad=pad, blend="crop")
for f in frames:
    ours = r.process_frame(f)
t0 = time.perf_counter()
for i in range(a.frames * 3):
    ours = r.process_frame(frames[i % 2])
ours_fps = a.frames * 3 / (time.perf_counter() - t0)
d = np.abs(ours.astype(np.int32) - enhance(frames[(a.frames * 3 - 1) % 2]).astype(np.int32))
print(json.dumps({"workload": a.workload, "what": "upscale stage only (tiled network, crop-merge), host frame in / host frame out, synchronous per frame",
                  "torch_cudnn_fp16_fps": round(torch_fps, 3), "libvrb200_fps": round(ours_fps, 3), "ratio": round(ours_fps / torch_fps, 2),
                  "max_lsb_between_the_two_fp16_paths": int(d.max()), "differing_frac": round(float((d > 0).mean()), 4),
                  "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "gpu": torch.cuda.get_device_name(0)}))
