"""Generates the committed golden fixtures (run in the build container, where OpenCV 4.13 is installed).

    python tests/golden/make_golden.py

Writes tests/golden/*.npz:
  filters_cv2.npz   input frame + cv2.bilateralFilter / cvtColor / createCLAHE outputs (the reference's own library,
                    video_upscaler.py:496; CLAHE per README.md:11,240) -- pins oracle/filters.py to OpenCV.
  filters_spec.npz  oracle outputs for the README-only stages (unsharp, temporal, blend window) -- pins the spec.
  tile_grids.npz    RealESRGANer.tile_process index tables for 24 (H,W,tile,pad,scale) cases.
  nets_<model>.npz  one 32x40 frame through each model (oracle fp32, random-init seed 0) -- regression pin of the
                    restatement; NOT a reference artefact (none exists: SURVEY.md section 4).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
OUT = Path(__file__).resolve().parent

TILE_CASES = [(256, 256, 128, 16, 4), (480, 854, 1024, 10, 4), (1080, 1920, 512, 32, 2), (720, 1280, 512, 64, 4),
              (720, 1280, 1536, 10, 4), (1080, 1920, 1024, 10, 4), (1080, 1920, 512, 32, 4), (37, 53, 16, 3, 4),
              (64, 64, 64, 0, 4), (65, 64, 64, 8, 4), (64, 65, 64, 8, 2), (1, 1, 8, 2, 4), (7, 300, 128, 10, 4),
              (300, 7, 128, 10, 4), (1080, 1920, 400, 10, 4), (1088, 1920, 512, 56, 4), (129, 129, 128, 1, 4),
              (128, 128, 128, 128, 4), (100, 100, 30, 40, 4), (2160, 3840, 1536, 32, 2), (480, 640, 512, 16, 4),
              (481, 641, 512, 16, 2), (33, 1000, 32, 5, 4), (1000, 33, 32, 5, 4)]


def main():
    import cv2

    from oracle import filters as F
    from oracle.pipeline import OracleRestorer
    from oracle.realesrganer import blend_window, tile_grid
    from util import oracle_model_from_sd
    from video_restore_b200.synth import random_state_dict, synth_frame

    cv2.setNumThreads(1)
    frame = synth_frame(96, 128, seed=7)
    frame2 = synth_frame(96, 128, seed=7, index=1)
    ragged = synth_frame(75, 101, seed=8)
    ycc = cv2.cvtColor(frame, cv2.COLOR_BGR2YCrCb)
    clahe = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8))
    ycc_r = cv2.cvtColor(ragged, cv2.COLOR_BGR2YCrCb)
    np.savez_compressed(
        OUT / "filters_cv2.npz", frame=frame, ragged=ragged,
        bilateral=cv2.bilateralFilter(frame, 5, 25, 25), bilateral_ragged=cv2.bilateralFilter(ragged, 5, 25, 25),
        ycrcb=ycc, bgr_back=cv2.cvtColor(ycc, cv2.COLOR_YCrCb2BGR),
        clahe_y=clahe.apply(np.ascontiguousarray(ycc[:, :, 0])),
        clahe_y_ragged=clahe.apply(np.ascontiguousarray(ycc_r[:, :, 0])),
        cv2_version=np.array(cv2.__version__))
    hist, lut, th, tw = F.clahe_tables(np.ascontiguousarray(ycc[:, :, 0]))
    np.savez_compressed(
        OUT / "filters_spec.npz", frame=frame, frame2=frame2, unsharp=F.unsharp_mask(frame, 0.5),
        temporal=F.temporal_blend(frame2, frame), clahe_bgr=F.clahe_bgr(frame), clahe_hist=hist, clahe_lut=lut,
        blend_window_96=blend_window(96), blend_window_2304=blend_window(2304), taps7=F.gaussian_taps7())
    np.savez_compressed(OUT / "tile_grids.npz", cases=np.asarray(TILE_CASES, np.int32),
                        **{f"grid_{i}": tile_grid(*c) for i, c in enumerate(TILE_CASES)})
    f = synth_frame(32, 40, seed=9)
    for name in ("RealESRGAN_x4plus", "RealESRGAN_x2plus", "RealESRGAN_x4plus_anime_6B", "RealESRGAN_x4_v3"):
        sd = random_state_dict(name, seed=0)
        orc = OracleRestorer(name, tile=24, tile_pad=4, model=oracle_model_from_sd(name, sd))
        np.savez_compressed(OUT / f"nets_{name}.npz", frame=f, out=orc.process_frame(f))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
