"""Small end-to-end run for compute-sanitizer: every kernel family once (x4plus anime 2x2 tiles, gaussian blend, all
filters, x2 model with odd extent, SRVGG)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from video_restore_b200.restorer import FrameOpts, FrameRestorer
from video_restore_b200.synth import random_state_dict, synth_frame
opts = FrameOpts(denoise=True, sharpen=0.4, clahe=True, temporal=True)
for name, tile, pad, blend, shape in [("RealESRGAN_x4plus_anime_6B", 48, 8, "gaussian", (70, 90)),
                                      ("RealESRGAN_x2plus", 32, 8, "crop", (65, 91)),
                                      ("RealESRGAN_x4_v3", 1024, 10, "crop", (64, 128))]:
    r = FrameRestorer(name, random_state_dict(name, 0), tile=tile, tile_pad=pad, blend=blend)
    for i in range(2):
        out = r.process_frame(synth_frame(*shape, seed=3, index=i), opts)
    print(name, out.shape, int(out.mean()), flush=True)
    r.close()
print("done")
