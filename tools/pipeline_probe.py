"""Where does the in-process pipeline lose time against a bare process_stream loop? python tools/pipeline_probe.py [frames]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from video_restore_b200.pipeline import ArraySource, NullSink, run_pipeline
from video_restore_b200.restorer import FrameOpts, FrameRestorer
from video_restore_b200.synth import random_state_dict, synth_frame

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
name = "RealESRGAN_x4plus"
sd = random_state_dict(name, 0)
frames = [synth_frame(720, 1280, seed=1, index=i) for i in range(8)]
src = ArraySource([frames[i % 8] for i in range(n)])
opts = FrameOpts()
mk = lambda g: FrameRestorer(name, sd, tile=1536, tile_pad=10, blend="crop", gpu_id=g)

r = mk(0)
for _ in r.process_stream(iter(frames[:2]), opts):
    pass
t0 = time.perf_counter()
k = 0
for out in r.process_stream(src.read_range(0, n), opts):
    k += 1
dt = time.perf_counter() - t0
print(f"[probe] bare process_stream: {k / dt:.2f} fps", flush=True)
r.close()

class Timed(NullSink):
    def __init__(self):
        super().__init__(); self.t = 0.0
    def write(self, i, f):
        t = time.perf_counter(); super().write(i, f); self.t += time.perf_counter() - t

for label, kw in (("pipeline zero-copy", {}), ("pipeline zero-copy again", {}), ("pipeline, no checksum sink", {"nosum": True}),
                  ("pipeline zero-copy, third time", {})):
    sink = Timed()
    if kw.get("nosum"):
        sink.write = lambda i, f, s=sink: (s.order.append(i), setattr(s, "count", s.count + 1))
    st = run_pipeline(src, sink, mk, [0], opts, chunk=None)
    print(f"[probe] {label}: {st.fps:.2f} fps ({st.frames} frames, {st.seconds:.2f} s, sink {sink.t:.3f} s, set-up {st.setup_seconds:.1f} s)", flush=True)
