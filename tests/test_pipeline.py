"""Multi-GPU in-process pipeline (video_restore_b200/pipeline.py): chunk plan, ordered reassembly with bounded memory,
boundary-frame hand-over. CPU tests drive the real scheduler with a stub restorer whose temporal stage is the oracle's;
the GPU test runs two real restorers (two host threads) and compares with one restorer walking the clip in order."""
import threading
import time

import numpy as np
import pytest

from oracle import filters as OF
from video_restore_b200.pipeline import (ArraySource, ListSink, NullSink, OrderedReassembler, SyntheticSource,
                                         plan_chunks, run_pipeline)
from video_restore_b200.restorer import FrameOpts


class StubRestorer:
    """x2 nearest 'upscale' + a per-frame tweak; temporal stage = oracle temporal_blend on the un-blended results."""
    scale = 2
    instances = []

    def __init__(self, gpu_id, delay=0.0):
        self.gpu_id, self.delay = gpu_id, delay
        self.prev = None
        self.closed = False
        self.frames_done = 0
        StubRestorer.instances.append(self)

    def _up(self, f):
        u = np.repeat(np.repeat(f, 2, axis=0), 2, axis=1).astype(np.int32)
        return np.clip(u + (u[::-1] % 7) - 3, 0, 255).astype(np.uint8)

    def temporal_reset(self):
        self.prev = None

    def temporal_get_prev(self, sH, sW):
        assert self.prev is not None and self.prev.shape[:2] == (sH, sW)
        return self.prev.copy()

    def process_stream(self, frames, opts):
        for f in frames:
            if self.delay:
                time.sleep(self.delay)
            u = self._up(f)
            out = u
            if opts.temporal:
                if self.prev is not None:
                    out = OF.temporal_blend(u, self.prev, opts.temporal_alpha, opts.temporal_tau)
                self.prev = u
            self.frames_done += 1
            yield out

    def close(self):
        self.closed = True


def stub_blend(gpu_id):
    return lambda cur, prev, alpha, tau: OF.temporal_blend(cur, prev, alpha, tau)


def clip(n, h=12, w=16, seed=3):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    # slow drift + occasional jumps so that the temporal gate both fires and does not
    return [np.clip(base.astype(np.int32) + (i % 5) * 2 + (40 if i % 7 == 0 else 0), 0, 255).astype(np.uint8) for i in range(n)]


def sequential(frames, opts):
    r = StubRestorer(-1)
    return list(r.process_stream(iter(frames), opts))


def test_plan_chunks_contiguous_and_chunked():
    assert plan_chunks(10, 3) == [(0, 4, 0), (4, 7, 1), (7, 10, 2)]
    assert plan_chunks(2, 4) == [(0, 1, 0), (1, 2, 1)]            # more GPUs than frames: empty ranges dropped
    assert plan_chunks(10, 2, 4) == [(0, 4, 0), (4, 8, 1), (8, 10, 0)]
    assert plan_chunks(0, 2, 4) == []
    for total, g, c in [(37, 3, 5), (64, 8, 8), (5, 2, 1)]:
        p = plan_chunks(total, g, c)
        assert p[0][0] == 0 and p[-1][1] == total and all(a[1] == b[0] for a, b in zip(p, p[1:]))
        assert all(w == i % g for i, (_, _, w) in enumerate(p))
    with pytest.raises(ValueError):
        plan_chunks(10, 0)
    with pytest.raises(ValueError):
        plan_chunks(10, 2, 0)


@pytest.mark.parametrize("n,gpus,chunk", [(23, [0, 1], 4), (23, [0, 1, 2], None), (9, [0, 1, 2, 3], 1), (16, [0], 5),
                                          (7, [0, 1, 2, 3, 4, 5, 6, 7], 2), (1, [0, 1], 3)])
def test_pipeline_matches_sequential_with_temporal(n, gpus, chunk):
    frames = clip(n)
    opts = FrameOpts(temporal=True)
    want = sequential(frames, opts)
    sink = ListSink()
    StubRestorer.instances.clear()
    warm = n % 2 == 0
    st = run_pipeline(ArraySource(frames), sink, lambda g: StubRestorer(g, delay=0.001 * (g % 3)), gpus, opts, chunk=chunk,
                      temporal_blend=stub_blend, warmup=warm)
    assert sink.order == list(range(n))
    assert len(sink.frames) == n and all(np.array_equal(a, b) for a, b in zip(sink.frames, want))
    assert st.frames == n and st.boundary_frames == st.chunks - 1
    # no redundant upscales: every frame went through exactly one restorer once (+ one warm-up frame per restorer)
    assert sum(r.frames_done for r in StubRestorer.instances) == n + (len(gpus) if warm else 0)
    assert all(r.closed for r in StubRestorer.instances)


def test_pipeline_without_temporal_has_no_boundary_traffic():
    frames = clip(13)
    opts = FrameOpts()
    sink = ListSink()
    st = run_pipeline(ArraySource(frames), sink, lambda g: StubRestorer(g), [0, 1, 2], opts, chunk=2, temporal_blend=stub_blend)
    assert st.boundary_frames == 0 and sink.order == list(range(13))
    assert all(np.array_equal(a, b) for a, b in zip(sink.frames, sequential(frames, opts)))


def test_reassembly_memory_is_bounded():
    """A fast worker cannot run more than `capacity` frames ahead of the writer."""
    n, G, C = 40, 2, 4
    frames = clip(n, 6, 8)

    class SlowSink(ListSink):
        def write(self, index, frame):
            time.sleep(0.002)
            super().write(index, frame)

    sink = SlowSink()
    st = run_pipeline(ArraySource(frames), sink, lambda g: StubRestorer(g), list(range(G)), FrameOpts(temporal=True), chunk=C,
                      temporal_blend=stub_blend)
    assert sink.order == list(range(n))
    assert st.max_held <= G * C + G
    with pytest.raises(ValueError):
        run_pipeline(ArraySource(frames), ListSink(), lambda g: StubRestorer(g), [0, 1], FrameOpts(temporal=True), chunk=4,
                     capacity=7, temporal_blend=stub_blend)
    # without deferred head frames any ring size works: a tiny ring only serialises the workers
    sink2 = ListSink()
    st2 = run_pipeline(ArraySource(frames), sink2, lambda g: StubRestorer(g), [0, 1], FrameOpts(), chunk=4, capacity=2)
    assert sink2.order == list(range(n)) and st2.max_held <= 2


def test_reassembler_blocks_and_orders():
    sink = ListSink()
    ra = OrderedReassembler(sink, total=6, capacity=2)
    f = np.zeros((2, 2, 3), np.uint8)
    blocked = threading.Event()

    def late():
        ra.put(3, f + 3)      # 3 >= 0 + 2: must block until frames 0 and 1 are written
        blocked.set()

    t = threading.Thread(target=late)
    t.start()
    time.sleep(0.05)
    assert not blocked.is_set()
    ra.put(1, f + 1)
    ra.put(0, f)
    ra.put(2, f + 2)
    t.join(2)
    assert blocked.is_set()
    ra.put(4, f + 4)
    ra.put(5, f + 5)
    ra.finish()
    assert sink.order == [0, 1, 2, 3, 4, 5] and [int(a[0, 0, 0]) for a in sink.frames] == [0, 1, 2, 3, 4, 5]


def test_worker_error_propagates_and_does_not_hang():
    class Boom(StubRestorer):
        def process_stream(self, frames, opts):
            for i, out in enumerate(super().process_stream(frames, opts)):
                if self.gpu_id == 1 and i == 1:
                    raise RuntimeError("boom")
                yield out

    with pytest.raises(RuntimeError, match="boom"):
        run_pipeline(ArraySource(clip(20)), NullSink(), lambda g: Boom(g), [0, 1], FrameOpts(temporal=True), chunk=3,
                     temporal_blend=stub_blend)


def test_synthetic_source_is_random_access():
    src = SyntheticSource(8, 12, 5, seed=2)
    a = list(src.reader().read_range(0, 5))
    b = list(src.reader().read_range(3, 5))
    assert len(src) == 5 and np.array_equal(a[3], b[0]) and np.array_equal(a[4], b[1])


@pytest.mark.gpu
def test_two_restorers_match_one_gpu():
    from video_restore_b200.restorer import FrameRestorer
    from video_restore_b200.synth import random_state_dict, synth_frame

    name = "RealESRGAN_x4_v3"
    sd = random_state_dict(name, seed=0)
    frames = [synth_frame(40, 56, seed=9, index=i) for i in range(11)]
    opts = FrameOpts(denoise=True, sharpen=0.3, clahe=True, temporal=True)
    one = FrameRestorer(name, sd, tile=64, tile_pad=10, gpu_id=0)
    one.temporal_reset()
    want = [one.process_frame(f, opts) for f in frames]
    one.close()
    for gpus, chunk in (([0, 0], 3), ([0, 0, 0], None)):
        sink = ListSink()
        st = run_pipeline(ArraySource(frames), sink, lambda g: FrameRestorer(name, sd, tile=64, tile_pad=10, gpu_id=g), gpus,
                          opts, chunk=chunk)
        assert sink.order == list(range(len(frames)))
        assert all(np.array_equal(a, b) for a, b in zip(sink.frames, want)), (gpus, chunk)
        assert st.boundary_frames == st.chunks - 1


def _write_clip(path, n, h=48, w=64):
    import cv2

    wr = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"mp4v"), 24.0, (w, h))
    assert wr.isOpened()
    for i in range(n):
        wr.write(np.full((h, w, 3), 20 + 25 * i, np.uint8))
    wr.release()


def test_video_file_source_and_sink_roundtrip(tmp_path):
    """cv2 decode -> chunked workers (each with its own VideoCapture, seeking to its chunk) -> ordered cv2 encode."""
    import cv2

    from video_restore_b200.pipeline import VideoFileSink, VideoFileSource

    src_path, dst_path = tmp_path / "in.mp4", tmp_path / "out.mp4"
    _write_clip(src_path, 9)
    src = VideoFileSource(str(src_path))
    assert len(src) == 9 and (src.height, src.width) == (48, 64) and abs(src.fps - 24.0) < 1e-3
    st = run_pipeline(src, VideoFileSink(str(dst_path), src.fps), lambda g: StubRestorer(g), [0, 1, 2], FrameOpts(), chunk=2,
                      temporal_blend=stub_blend)
    assert st.frames == 9
    cap = cv2.VideoCapture(str(dst_path))
    means = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        assert f.shape == (96, 128, 3)
        means.append(float(f.mean()))
    assert len(means) == 9
    assert all(b > a for a, b in zip(means, means[1:])), "frames must come out in source order"
    with pytest.raises(OSError):
        VideoFileSource(str(tmp_path / "missing.mp4"))
    # both ways of reaching a chunk's first frame deliver the same frames
    a = list(VideoFileSource(str(src_path), seek="set").reader().read_range(5, 8))
    rd = VideoFileSource(str(src_path), seek="grab").reader()
    b = list(rd.read_range(5, 8))
    assert len(a) == len(b) == 3 and all(np.array_equal(x, y) for x, y in zip(a, b))
    c = list(rd.read_range(2, 4))  # going back re-opens the file
    assert len(c) == 2 and np.array_equal(c[0], list(VideoFileSource(str(src_path)).reader().read_range(2, 3))[0])


def test_video_file_source_counts_exactly_and_decodes_sequentially(tmp_path):
    """the container's frame count is a hint. The source counts frames itself (grab pass), the default reader is ONE
    sequential decoder shared by all workers (no seeks), frames past EOF end the range, and a frame is handed out once."""
    from video_restore_b200.pipeline import ListSink, VideoFileSource

    p = tmp_path / "c.mp4"
    _write_clip(p, 10)
    src = VideoFileSource(str(p), lookahead=4)          # look-ahead smaller than G * chunk: must not deadlock
    assert len(src) == 10
    sink = ListSink()
    st = run_pipeline(src, sink, lambda g: StubRestorer(g), [0, 1, 2], FrameOpts(temporal=True), chunk=2,
                      temporal_blend=stub_blend)
    assert st.frames == 10 and sink.order == list(range(10))
    src = VideoFileSource(str(p))
    rd = src.reader()
    assert len(list(rd.read_range(8, 20))) == 2         # stops at EOF
    assert len(list(rd.read_range(3, 4))) == 1          # still buffered: decoded on the way to frame 8, never taken
    with pytest.raises(RuntimeError, match="already consumed"):
        list(rd.read_range(3, 4))
    assert len(list(rd.read_range(0, 1))) == 1          # frame 0 stays available (warm-up reads)
    src.close()
    # one long range per worker: per-worker exact readers instead of the shared decoder
    src = VideoFileSource(str(p))
    sink2 = ListSink()
    run_pipeline(src, sink2, lambda g: StubRestorer(g), [0, 1], FrameOpts(temporal=True), chunk=None, temporal_blend=stub_blend)
    assert sink2.order == list(range(10))
    assert all(np.array_equal(a, b) for a, b in zip(sink.frames, sink2.frames))


@pytest.mark.gpu
def test_cli_video_file_two_workers(tmp_path, capsys):
    import cv2

    from video_restore_b200.cli import main

    src_path, dst_path = tmp_path / "in.mp4", tmp_path / "out.mp4"
    _write_clip(src_path, 7)
    assert main([str(src_path), str(dst_path), "--model", "RealESRGAN_x4_v3", "--quality", "fast", "--enhanced",
                 "--gpus", "0", "0", "--random-weights"]) == 0
    assert "processed 7 frames" in capsys.readouterr().out
    cap = cv2.VideoCapture(str(dst_path))
    n = 0
    while True:
        ok, f = cap.read()
        if not ok:
            break
        assert f.shape == (192, 256, 3)
        n += 1
    assert n == 7
    # directory mode
    out_dir = tmp_path / "outs"
    assert main([str(tmp_path), str(out_dir), "--model", "RealESRGAN_x4_v3", "--quality", "fast", "--batch", "--gpus", "0",
                 "--random-weights"]) == 0
    assert (out_dir / "in_upscaled.mp4").exists()


@pytest.mark.gpu
def test_video_file_content_matches_oracle(tmp_path):
    """N1 content, not only counts: frames decoded by VideoFileSource -> two pipeline workers -> (a) an in-memory sink compared
    with the CPU oracle applied to the SAME decoded frames (upscale +-1 LSB, the bit-exact filters composed on top of our
    upscale), (b) the encoded file re-read and compared with (a) within the codec's loss."""
    import cv2

    from oracle.pipeline import OracleRestorer
    from video_restore_b200.pipeline import VideoFileSink, VideoFileSource
    from video_restore_b200.restorer import FrameRestorer
    from video_restore_b200.synth import random_state_dict, synth_frame

    from util import oracle_model_from_sd

    src_path, dst_path = tmp_path / "in.mp4", tmp_path / "out.mp4"
    wr = cv2.VideoWriter(str(src_path), cv2.VideoWriter_fourcc(*"mp4v"), 24.0, (64, 48))
    assert wr.isOpened()
    for i in range(6):
        wr.write(synth_frame(48, 64, seed=4, index=i))       # textured, moving content
    wr.release()
    cap = cv2.VideoCapture(str(src_path))
    decoded = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        decoded.append(f)
    assert len(decoded) == 6 and decoded[0].std() > 5.0
    name = "RealESRGAN_x4_v3"
    sd = random_state_dict(name, seed=0)
    plain, enh = FrameOpts(), FrameOpts(sharpen=0.3, clahe=True)
    results = {}
    for key, opts in (("plain", plain), ("enh", enh)):
        src = VideoFileSource(str(src_path))
        sink = ListSink()
        st = run_pipeline(src, sink, lambda g: FrameRestorer(name, sd, tile=32, tile_pad=8, gpu_id=g), [0, 0], opts, chunk=2)
        assert st.frames == 6 and sink.order == list(range(6))
        results[key] = sink.frames
    up = OracleRestorer(name, tile=32, tile_pad=8, model=oracle_model_from_sd(name, sd))
    for got, got_enh, f in zip(results["plain"], results["enh"], decoded):
        d = np.abs(got.astype(np.int32) - up.process_frame(f, plain).astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 0.05
        # the enhancement filters are bit-exact given their input (CLAHE can amplify the upscale's one level into a few, so the
        # enhanced frame is checked compositionally: the oracle's filters on OUR upscale)
        assert np.array_equal(got_enh, OF.clahe_bgr(OF.unsharp_mask(got, 0.3)))
    # (b) through the encoder: mp4v is lossy -- the re-read file must be the same pictures within the codec's error
    src = VideoFileSource(str(src_path))
    run_pipeline(src, VideoFileSink(str(dst_path), src.fps), lambda g: FrameRestorer(name, sd, tile=32, tile_pad=8, gpu_id=g), [0, 0],
                 enh, chunk=2)
    cap = cv2.VideoCapture(str(dst_path))
    for want in results["enh"]:
        ok, f = cap.read()
        assert ok and f.shape == want.shape
        mse = np.mean((f.astype(np.float64) - want.astype(np.float64)) ** 2)
        assert 10 * np.log10(255.0 ** 2 / max(mse, 1e-12)) > 20.0  # unrelated pictures: ~10 dB
    assert not cap.read()[0]


class StubZeroCopy(StubRestorer):
    """Same arithmetic, but renders into the pipeline's buffer pool like FrameRestorer.process_stream(out_pool=...)."""
    zero_copy_stream = True
    allocs = 0

    @staticmethod
    def alloc_host(shape):
        StubZeroCopy.allocs += 1
        return np.empty(shape, np.uint8)

    def temporal_get_prev(self, sH, sW, out=None):
        if out is None:
            return super().temporal_get_prev(sH, sW)
        np.copyto(out, self.prev)
        return out

    def process_stream(self, frames, opts, out_pool=None):
        for out in super().process_stream(frames, opts):
            if out_pool is None:
                yield out
            else:
                dst = out_pool.get(out.shape)
                np.copyto(dst, out)
                yield dst


@pytest.mark.parametrize("gpus,chunk,temporal", [([0, 1], 4, True), ([0, 1, 2], 3, False), ([0], None, True)])
def test_zero_copy_pool_matches_and_is_bounded(gpus, chunk, temporal):
    n = 41
    frames = clip(n, 6, 8)
    opts = FrameOpts(temporal=temporal)
    want = sequential(frames, opts)

    class SlowSink(ListSink):
        def write(self, index, frame):
            time.sleep(0.001)
            super().write(index, frame)

    sink = SlowSink()
    StubZeroCopy.allocs = 0
    st = run_pipeline(ArraySource(frames), sink, lambda g: StubZeroCopy(g), gpus, opts, chunk=chunk, temporal_blend=stub_blend)
    assert sink.order == list(range(n))
    assert all(np.array_equal(a, b) for a, b in zip(sink.frames, want))
    G = len(gpus)
    C = chunk or n
    cap = (G * C + G) if (temporal and G > 1) else max(min(G * C, 64), 4)
    assert StubZeroCopy.allocs <= cap + 5 * G, "host frame memory must stay within the pool"


def test_source_shorter_than_announced_does_not_hang():
    """Container frame counts are estimates: a source that runs dry early truncates the output instead of stalling."""
    class Liar(ArraySource):
        def __len__(self):
            return len(self.frames) + 5

    frames = clip(14)
    opts = FrameOpts(temporal=True)
    want = sequential(frames, opts)
    for gpus, chunk in (([0, 1], 4), ([0], None), ([0, 1, 2], 2)):
        sink = ListSink()
        st = run_pipeline(Liar(frames), sink, lambda g: StubZeroCopy(g), gpus, opts, chunk=chunk, temporal_blend=stub_blend)
        assert st.frames == 14 and sink.order == list(range(14))
        assert all(np.array_equal(a, b) for a, b in zip(sink.frames, want))


def test_restorer_construction_failure_propagates():
    def make(g):
        if g == 1:
            raise ValueError("no such device")
        return StubRestorer(g)

    with pytest.raises(ValueError, match="no such device"):
        run_pipeline(ArraySource(clip(10)), NullSink(), make, [0, 1, 2], FrameOpts(temporal=True), chunk=2, temporal_blend=stub_blend)


def test_buffer_pool_byte_cap():
    """the pinned pool is capped by BYTES once the ring's minimum exists, not only by frame count."""
    from video_restore_b200.pipeline import BufferPool

    made = []
    pool = BufferPool(count=10, alloc=lambda s: (made.append(s), np.empty(s, np.uint8))[1], min_count=3, max_bytes=4 * 100)
    bufs = [pool.get((10, 10)) for _ in range(3)]       # the minimum is always granted (3 x 100 bytes)
    bufs.append(pool.get((10, 10)))                     # 4 x 100 <= cap
    assert len(made) == 4

    got = []
    t = threading.Thread(target=lambda: got.append(pool.get((10, 10))), daemon=True)   # a fifth would exceed the cap: waits
    t.start()
    time.sleep(0.3)
    assert not got and len(made) == 4
    pool.release(bufs.pop())
    t.join(2.0)
    assert got and len(made) == 4                       # it got the released buffer, nothing new was pinned


# --- the reference's ffmpeg rawvideo pipes (video_upscaler.py:220-262, :514-532) against stand-in binaries ---------------
_FAKE_FFPROBE = '''#!{py}
import sys, json, numpy as np
a = np.load(sys.argv[-1])
assert "-count_frames" in sys.argv and "v:0" in sys.argv
print(json.dumps({{"streams": [{{"width": a.shape[2], "height": a.shape[1], "r_frame_rate": "24000/1001", "nb_read_frames": str(a.shape[0])}}]}}))
'''
_FAKE_FFMPEG = '''#!{py}
import sys, json, numpy as np
argv = sys.argv[1:]
if "-hwaccels" in argv:
    print("Hardware acceleration methods:\\nvdpau\\n{hw}")
    sys.exit(0)
src = argv[argv.index("-i") + 1]
if src == "-":                                   # encoder: raw BGR frames on stdin -> the "file"
    data = sys.stdin.buffer.read()
    open(argv[-1], "wb").write(data)
    open(argv[-1] + ".argv.json", "w").write(json.dumps(argv))
    sys.exit({enc_rc})
assert argv[-1] == "-" and argv[argv.index("-pix_fmt") + 1] == "bgr24" and argv[argv.index("-f") + 1] == "rawvideo"
open(src + ".decode_argv.json", "w").write(json.dumps(argv))
try:                                             # decoder: every frame of the .npy "video" to stdout
    sys.stdout.buffer.write(np.load(src).tobytes())
    sys.stdout.buffer.flush()
except BrokenPipeError:
    pass
'''


def _fake_ffmpeg_tools(tmp_path, hw="", enc_rc=0):
    import stat
    import sys

    d = tmp_path / "bin"
    d.mkdir(exist_ok=True)
    for name, body in (("ffprobe", _FAKE_FFPROBE.format(py=sys.executable)),
                       ("ffmpeg", _FAKE_FFMPEG.format(py=sys.executable, hw=hw, enc_rc=enc_rc))):
        f = d / name
        f.write_text(body)
        f.chmod(f.stat().st_mode | stat.S_IEXEC)
    return str(d / "ffmpeg"), str(d / "ffprobe")


def test_ffmpeg_pipe_source_counts_and_reads_exactly(tmp_path):
    from video_restore_b200.pipeline import FfmpegPipeSource

    ffmpeg, ffprobe = _fake_ffmpeg_tools(tmp_path, hw="cuda")
    frames = clip(11)
    video = tmp_path / "in.npy"
    np.save(video, np.stack(frames))
    src = FfmpegPipeSource(str(video), ffmpeg, ffprobe, lookahead=4)
    assert len(src) == 11 and (src.height, src.width) == (12, 16) and abs(src.fps - 24000 / 1001) < 1e-9
    assert src.hwaccel == "cuda"                                          # listed by `ffmpeg -hwaccels` (:264-278)
    # interleaved chunks through the shared front-to-back decoder
    r0, r1 = src.reader(), src.reader()
    got = {}
    for rd, (a, b) in ((r0, (0, 3)), (r1, (3, 6)), (r0, (6, 9)), (r1, (9, 14))):
        for i, f in zip(range(a, b), rd.read_range(a, b)):
            got[i] = f
    assert sorted(got) == list(range(11)) and all(np.array_equal(got[i], frames[i]) for i in got)   # EOF ends the last range
    src.close()
    import json

    argv = json.loads((tmp_path / "in.npy.decode_argv.json").read_text())
    assert argv.index("-hwaccel") < argv.index("-i") and argv[argv.index("-hwaccel") + 1] == "cuda"  # before the input (:227)
    # one long range per worker: a decoder process of its own, exact forward skipping, going back re-opens
    rd = src.reader(long_ranges=True)
    assert all(np.array_equal(a, b) for a, b in zip(rd.read_range(7, 10), frames[7:10]))
    assert all(np.array_equal(a, b) for a, b in zip(rd.read_range(2, 4), frames[2:4]))
    rd.cap.release()
    with pytest.raises(RuntimeError, match="Failed to read video info"):
        FfmpegPipeSource(str(tmp_path / "missing.npy"), ffmpeg, ffprobe)


def test_ffmpeg_pipe_sink_and_pipeline_roundtrip(tmp_path):
    """ffmpeg decode pipe -> three chunked workers -> ordered reassembly -> ffmpeg libx264 encode pipe (stand-in binaries: the
    'encoded file' is the raw byte stream the encoder was fed)."""
    import json

    from video_restore_b200.pipeline import FfmpegPipeSink, FfmpegPipeSource

    ffmpeg, ffprobe = _fake_ffmpeg_tools(tmp_path)
    frames = clip(13)
    video, out = tmp_path / "in.npy", tmp_path / "out.mp4"
    np.save(video, np.stack(frames))
    opts = FrameOpts(temporal=True)
    src = FfmpegPipeSource(str(video), ffmpeg, ffprobe)
    assert src.hwaccel is None
    st = run_pipeline(src, FfmpegPipeSink(str(out), src.fps, crf=12, preset="veryslow", ffmpeg_bin=ffmpeg), lambda g: StubRestorer(g),
                      [0, 1, 2], opts, chunk=2, temporal_blend=stub_blend)
    assert st.frames == 13
    want = np.stack(sequential(frames, opts))
    assert out.read_bytes() == want.tobytes()                             # every frame, in order, bit for bit
    argv = json.loads((tmp_path / "out.mp4.argv.json").read_text())
    for flag, val in (("-s", "32x24"), ("-vcodec", "rawvideo"), ("-crf", "12"), ("-preset", "veryslow"), ("-r", str(src.fps)),
                      ("-movflags", "+faststart")):
        assert argv[argv.index(flag) + 1] == val, flag
    assert "libx264" in argv and "-an" in argv and argv.count("-pix_fmt") == 2 and argv[-1] == str(out)
    # an encoder that fails is an error, not a silently short file
    bad, _ = _fake_ffmpeg_tools(tmp_path, enc_rc=3)
    sink = FfmpegPipeSink(str(out), 24.0, ffmpeg_bin=bad)
    sink.write(0, want[0])
    with pytest.raises(OSError, match="exited with 3"):
        sink.close()


def test_video_io_selection(tmp_path, monkeypatch):
    from video_restore_b200 import pipeline as P

    ffmpeg, _ = _fake_ffmpeg_tools(tmp_path)
    video = tmp_path / "in.npy"
    np.save(video, np.stack(clip(3)))
    monkeypatch.setenv("PATH", str(tmp_path / "bin"))
    assert P.find_ffmpeg() == (ffmpeg, str(tmp_path / "bin" / "ffprobe"))
    assert isinstance(P.open_video_source(str(video)), P.FfmpegPipeSource)
    assert isinstance(P.open_video_sink(str(tmp_path / "o.mp4"), 24.0, 18, "fast"), P.FfmpegPipeSink)
    monkeypatch.setenv("VR_IO", "cv2")                                    # force OpenCV
    assert P.find_ffmpeg() is None
    assert isinstance(P.open_video_sink(str(tmp_path / "o.mp4"), 24.0), P.VideoFileSink)
    monkeypatch.delenv("VR_IO")
    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))                 # no binaries: OpenCV
    assert P.find_ffmpeg() is None
