"""GPU bring-up probe for the whole path: CUDA restore vs the CPU oracle on small frames (run with gpurun)."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import torch
from util import oracle_model_from_sd, psnr_u8, max_lsb, random_state_dict, synth_frame
from oracle.pipeline import OracleRestorer, FrameOpts as OOpts
from oracle import filters as OF
from oracle.realesrganer import tile_grid as o_tile_grid, blend_window
from video_restore_b200 import restorer as R

torch.set_num_threads(8)

def net_case(name, H, W, tile, pad, blend="crop", opts=None, nframes=1):
    sd = random_state_dict(name, seed=0)
    om = oracle_model_from_sd(name, sd)
    orc = OracleRestorer(name, tile=tile, tile_pad=pad, blend=blend, model=om)
    gpu = R.FrameRestorer(name, sd, tile=tile, tile_pad=pad, blend=blend)
    for i in range(nframes):
        f = synth_frame(H, W, seed=3, index=i)
        t0 = time.time(); ref = orc.process_frame(f, OOpts(**(opts or {}))); t1 = time.time()
        out = gpu.process_frame(f, R.FrameOpts(**(opts or {}))); t2 = time.time()
        d = np.abs(ref.astype(int) - out.astype(int))
        print(f"[{name} {H}x{W} tile={tile}/{pad} {blend} {opts} f{i}] max={d.max()} frac>0={np.mean(d>0):.4f} "
              f"frac>1={np.mean(d>1):.5f} psnr={psnr_u8(ref,out):.2f} ref_mean={ref.mean():.1f} ref_std={ref.std():.1f} "
              f"cpu={t1-t0:.2f}s gpu={t2-t1:.3f}s timing={gpu.last_timing()}", flush=True)
    gpu.close()

g = sys.argv[1] if len(sys.argv) > 1 else "all"
if g in ("all", "filters"):
    f = synth_frame(97, 131, seed=1)
    f2 = synth_frame(97, 131, seed=1, index=1)
    print("tile_grid eq:", all(np.array_equal(R.tile_grid(h, w, t, p, s), o_tile_grid(h, w, t, p, s))
          for (h, w, t, p, s) in [(256,256,128,16,4),(1080,1920,512,32,2),(720,1280,512,64,4),(37,53,16,3,4),(480,854,1024,10,4)]))
    print("bilateral diff:", max_lsb(R.bilateral_filter(f), OF.bilateral_filter(f)), np.mean(R.bilateral_filter(f) != OF.bilateral_filter(f)))
    print("unsharp diff:", max_lsb(R.unsharp_mask(f, 0.5), OF.unsharp_mask(f, 0.5)))
    o, hist, lut = R.clahe_bgr(f, return_tables=True)
    yc = OF.bgr_to_ycrcb(f)
    oh, ol, _, _ = OF.clahe_tables(np.ascontiguousarray(yc[:, :, 0]))
    print("clahe hist eq:", np.array_equal(hist, oh), "lut eq:", np.array_equal(lut, ol), "img diff:", max_lsb(o, OF.clahe_bgr(f)))
    print("temporal diff:", max_lsb(R.temporal_blend(f2, f), OF.temporal_blend(f2, f)), "blended frac", np.mean(OF.temporal_blend(f2, f) != f2))
    for e in (96, 1000):
        print("blend w maxabs:", np.abs(R.blend_weights(e) - blend_window(e)).max())
if g in ("all", "nets"):
    net_case("RealESRGAN_x4_v3", 60, 100, 1024, 10)
    net_case("RealESRGAN_x4plus_anime_6B", 48, 64, 1024, 10)
    net_case("RealESRGAN_x4plus", 64, 64, 48, 8)
    net_case("RealESRGAN_x2plus", 66, 90, 32, 8)
    net_case("RealESRGAN_x2plus", 65, 91, 32, 8, blend="gaussian")
    net_case("RealESRGAN_x4plus_anime_6B", 70, 90, 32, 8, blend="gaussian",
             opts=dict(denoise=True, sharpen=0.5, clahe=True, temporal=True), nframes=3)
