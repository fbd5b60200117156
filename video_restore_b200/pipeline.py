"""In-process multi-GPU video pipeline: frame source -> per-GPU restorers -> ordered reassembly -> sink.

Replaces, for the restoration stage only, the reference's reader thread / per-GPU worker threads / writer thread
(video_upscaler.py:369-404 start-up, :406-451 reader, :453-488 workers, :507-567 writer) -- SURVEY.md 8(f) N1, N2:

* threading model as the reference: ONE host thread per GPU, each with its own restorer (`self.models[gpu_id]`, :340);
  ctypes releases the GIL inside the C-ABI calls, so the GPUs run concurrently;
* work split: CONTIGUOUS frame ranges ("chunks") instead of the reference's `frame_idx % n_gpus` tag on a shared queue
  (:437, which drops frames dequeued by the wrong worker, :471-473). Chunk c = frames [c*C, (c+1)*C) goes to GPU c % G.
  With C = ceil(F / G) this is exactly the one-range-per-GPU sharding of `sharder.py`; a sequential encoder needs the
  frames in order, so for it C is small and the ranges interleave;
* temporal consistency across a chunk boundary WITHOUT any redundant upscale: out_t = blend(u_t, u_{t-1}) depends on
  the previous frame's UN-blended result only (non-recursive, oracle/filters.py::temporal_blend). A worker therefore
  runs its chunk with a reset temporal state -- the head frame passes through as u_first -- publishes its last
  un-blended frame u_last for the next chunk, and finishes its own head frame with one stand-alone temporal kernel
  once the previous chunk's u_last has arrived. Exactly one boundary frame per chunk crosses GPUs; results are
  bit-identical to a single-GPU run (tests/test_pipeline.py);
* ordered reassembly with bounded memory (the reference's PriorityQueue + sentinel/timeout logic, :535-567): workers
  `put(i, frame)` into a ring of `capacity` frames and block while `i >= next_to_write + capacity`; one writer thread
  hands frames to the sink strictly in order. capacity >= G*C is required (a chunk's head frame is written last) and
  enforced.

No CPU fallback: restorers are `FrameRestorer` instances (CUDA only); tests drive the same scheduler with a stub.
"""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np


# ---------------------------------------------------------------------------------------------------------------
# sources and sinks
# ---------------------------------------------------------------------------------------------------------------
class SyntheticSource:
    """`n` generated frames (synth.synth_frame); random access, thread-safe. `distinct` > 0 generates only that many
    different frames and cycles through them (frame generation is ~50 ms of numpy per 720p frame, more than a B200
    needs to restore one)."""

    def __init__(self, height: int, width: int, n: int, seed: int = 1, distinct: int = 0):
        self.height, self.width, self.n, self.seed, self.distinct = height, width, n, seed, distinct
        self.fps = 30.0
        self._cache: dict = {}
        self._lock = threading.Lock()

    def __len__(self) -> int:
        return self.n

    def reader(self) -> "SyntheticSource":
        return self

    def _frame(self, i: int) -> np.ndarray:
        from .synth import synth_frame

        if self.distinct <= 0:
            return synth_frame(self.height, self.width, seed=self.seed, index=i)
        k = i % self.distinct
        with self._lock:
            if k not in self._cache:
                self._cache[k] = synth_frame(self.height, self.width, seed=self.seed, index=k)
            return self._cache[k]

    def read_range(self, start: int, end: int):
        for i in range(start, end):
            yield self._frame(i)


class ArraySource:
    """Frames held in memory (tests, small clips)."""

    def __init__(self, frames: Sequence[np.ndarray], fps: float = 30.0):
        self.frames, self.fps = list(frames), fps

    def __len__(self) -> int:
        return len(self.frames)

    def reader(self) -> "ArraySource":
        return self

    def read_range(self, start: int, end: int):
        return iter(self.frames[start:end])


class VideoFileSource:
    """OpenCV-decoded video file (stands in for the reference's ffmpeg rawvideo pipe, :220-249).

    Frame count: container metadata is a HINT (0, negative, too small or too large all occur; the reference falls back to
    `ffprobe -count_frames`, video_upscaler.py:195-203, and to "decode until EOF", :450, :540). Unless `trust_count=True` the
    frames are counted exactly up front with one grab() pass (no colour conversion; far cheaper than restoring them).

    Getting to the start of a chunk, `seek`:
      "sequential" (default) -- ONE decoder thread reads the file front to back, like the reference's single decode thread
          (:430-451), and hands every frame to whichever worker asks for its index (bounded look-ahead): never seeks, exact on
          any stream, decodes every frame once. Needs interleaved chunks (the CLI's 16-frame chunks); with one long range per
          worker the readers fall back to "grab".
      "grab" -- one decoder per worker that never seeks: skips forward with grab() (exact, decodes skipped frames).
      "set"  -- one decoder per worker, CAP_PROP_POS_FRAMES seeks (fast; OpenCV's backends may land on a neighbouring frame in
          inter-coded streams, so duplicated / dropped frames at chunk boundaries are possible: opt-in only)."""

    def __init__(self, path: str, seek: str = "sequential", trust_count: bool = False, lookahead: int = 256):
        import cv2

        if seek not in ("sequential", "set", "grab"):
            raise ValueError("seek must be 'sequential', 'grab' or 'set'")
        self.path, self.seek, self.lookahead = str(path), seek, int(lookahead)
        cap = cv2.VideoCapture(self.path)
        if not cap.isOpened():
            raise OSError(f"cannot open {self.path}")
        hint = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        self.fps = float(cap.get(cv2.CAP_PROP_FPS) or 30.0)
        self.width = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
        self.height = int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
        if trust_count and hint > 0:
            self.n = hint
        else:
            n = 0
            while cap.grab():
                n += 1
            self.n = n
        self.count_hint = hint
        cap.release()
        self._dispatch = None
        self._lock = threading.Lock()

    def __len__(self) -> int:
        return self.n

    def reader(self, long_ranges: bool = False):
        """A reader for one worker thread. `long_ranges`: the caller walks one long contiguous range per worker, which a
        shared sequential decoder would serialise."""
        if self.seek == "sequential" and not long_ranges:
            with self._lock:
                if self._dispatch is None:
                    self._dispatch = _SequentialDecoder(self.path, self.lookahead)
            return _SharedReader(self._dispatch)
        return _VideoReader(self.path, "grab" if self.seek == "sequential" else self.seek)

    def close(self) -> None:
        if self._dispatch is not None:
            self._dispatch.stop()
            self._dispatch = None


class _SequentialDecoder:
    """One thread decoding the file front to back into a bounded index -> frame map; consumers take frames by index (each
    frame once; frame 0 stays available for the workers' warm-up reads)."""

    def __init__(self, path: str, lookahead: int, open_capture: Optional[Callable[[], object]] = None):
        if open_capture is None:
            import cv2

            open_capture = lambda: cv2.VideoCapture(path)  # noqa: E731
        self._cap = open_capture()
        self._cv = threading.Condition()
        self._frames: dict = {}
        self._first = None
        self._next = 0            # next index to decode
        self._eof: Optional[int] = None
        self._stop = False
        self._lookahead = max(lookahead, 4)
        self._waiting: set = set()  # indices consumers are blocked on
        self._thread = threading.Thread(target=self._run, name="vr-decode", daemon=True)
        self._thread.start()

    def _run(self) -> None:
        while True:
            with self._cv:
                # bounded look-ahead, except that a frame somebody is waiting for is always decoded (no deadlock on a full map)
                while (not self._stop and len(self._frames) >= self._lookahead
                       and not any(w >= self._next for w in self._waiting)):
                    self._cv.wait(0.2)
                if self._stop:
                    break
            ok, frame = self._cap.read()
            with self._cv:
                if not ok:
                    self._eof = self._next
                    self._cv.notify_all()
                    break
                if self._next == 0:
                    self._first = frame
                self._frames[self._next] = frame
                self._next += 1
                self._cv.notify_all()
        self._cap.release()

    def take(self, i: int):
        with self._cv:
            if i == 0 and self._first is not None:
                self._frames.pop(0, None)
                return self._first
            self._waiting.add(i)
            try:
                while i not in self._frames:
                    if i == 0 and self._first is not None:
                        return self._first
                    if self._eof is not None and i >= self._eof:
                        return None
                    if i < self._next:
                        raise RuntimeError(f"frame {i} was already consumed (the sequential decoder hands every frame out once)")
                    if self._stop:
                        return None
                    self._cv.notify_all()
                    self._cv.wait(0.2)
                return self._frames.pop(i)
            finally:
                self._waiting.discard(i)
                self._cv.notify_all()

    def stop(self) -> None:
        with self._cv:
            self._stop = True
            self._cv.notify_all()


class _SharedReader:
    def __init__(self, dec: _SequentialDecoder):
        self._dec = dec

    def read_range(self, start: int, end: int):
        for i in range(start, end):
            f = self._dec.take(i)
            if f is None:
                return
            yield f


class _VideoReader:
    def __init__(self, path: str, seek: str = "grab", open_capture: Optional[Callable[[], object]] = None):
        if open_capture is None:
            import cv2

            self._cv2 = cv2
            open_capture = lambda: cv2.VideoCapture(path)  # noqa: E731
        elif seek == "set":
            raise ValueError("a pipe cannot seek: use seek='grab'")
        self._open = open_capture
        self.path, self.seek = path, seek
        self.cap = open_capture()
        self.pos = 0

    def read_range(self, start: int, end: int):
        if self.pos != start:
            if self.seek == "set":
                self.cap.set(self._cv2.CAP_PROP_POS_FRAMES, start)
                self.pos = start
            else:
                if self.pos > start:  # chunks come in increasing order per worker; re-open if a caller goes back
                    self.cap.release()
                    self.cap = self._open()
                    self.pos = 0
                while self.pos < start:
                    if not self.cap.grab():
                        return
                    self.pos += 1
        for _ in range(start, end):
            ok, frame = self.cap.read()
            if not ok:
                return
            self.pos += 1
            yield frame


# ---------------------------------------------------------------------------------------------------------------
# ffmpeg pipes: the reference's own decode / encode hand-off (video_upscaler.py:220-262, :514-532), used when the binaries exist
# ---------------------------------------------------------------------------------------------------------------
def find_ffmpeg() -> Optional[tuple]:
    """(ffmpeg, ffprobe) when both binaries are on PATH and VR_IO is not 'cv2', else None."""
    import os
    import shutil

    if os.environ.get("VR_IO", "").lower() == "cv2":
        return None
    a, b = shutil.which("ffmpeg"), shutil.which("ffprobe")
    return (a, b) if a and b else None


class _PipeCapture:
    """The `read` / `grab` / `release` part of cv2.VideoCapture over `ffmpeg -i <path> -f rawvideo -pix_fmt bgr24 -`: the
    decoder process writes whole BGR frames to its stdout, a short read is the end of the stream (:243-247)."""

    def __init__(self, exe: str, path: str, width: int, height: int, hwaccel: Optional[str] = None):
        import subprocess

        cmd = [exe, "-loglevel", "error"]
        if hwaccel:  # must precede the input (:227-229)
            cmd += ["-hwaccel", hwaccel]
        cmd += ["-i", path, "-f", "rawvideo", "-pix_fmt", "bgr24", "-"]
        self.shape = (height, width, 3)
        self.nbytes = height * width * 3
        self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, bufsize=0)

    def _read_exact(self) -> Optional[bytearray]:
        buf = bytearray(self.nbytes)
        view, got = memoryview(buf), 0
        while got < self.nbytes:
            n = self.proc.stdout.readinto(view[got:])
            if not n:
                return None
            got += n
        return buf

    def read(self):
        if self.proc is None:
            return False, None
        buf = self._read_exact()
        if buf is None:
            return False, None
        return True, np.frombuffer(buf, np.uint8).reshape(self.shape)

    def grab(self) -> bool:
        return self.read()[0]

    def release(self) -> None:
        if self.proc is not None:
            self.proc.stdout.close()
            self.proc.terminate()
            self.proc.wait()
            self.proc = None


class FfmpegPipeSource:
    """A video file decoded by the `ffmpeg` binary through a rawvideo pipe, as the reference does. Stream facts and the EXACT
    frame count come from one `ffprobe -count_frames` call (the reference's last-resort counter, :195-203, used first here:
    the chunk plan needs the true length). `-hwaccel cuda` / `nvdec` is requested when `ffmpeg -hwaccels` lists it (:264-278).
    Same reader contract as VideoFileSource: one shared front-to-back decoder for interleaved chunks, one forward-skipping
    decoder process per worker for long ranges."""

    def __init__(self, path: str, ffmpeg_bin: Optional[str] = None, ffprobe_bin: Optional[str] = None, lookahead: int = 256):
        import json
        import subprocess

        found = find_ffmpeg()
        self.ffmpeg = ffmpeg_bin or (found[0] if found else None)
        self.ffprobe = ffprobe_bin or (found[1] if found else None)
        if not self.ffmpeg or not self.ffprobe:
            raise OSError("ffmpeg / ffprobe not found")
        self.path, self.lookahead = str(path), int(lookahead)
        done = subprocess.run([self.ffprobe, "-v", "error", "-select_streams", "v:0", "-count_frames", "-show_entries",
                               "stream=width,height,r_frame_rate,nb_read_frames", "-of", "json", self.path],
                              capture_output=True, text=True, check=False)
        try:
            st = json.loads(done.stdout)["streams"][0]
            self.width, self.height = int(st["width"]), int(st["height"])
            num, _, den = str(st["r_frame_rate"]).partition("/")
            self.fps = float(num) / float(den or 1)
            self.n = int(st["nb_read_frames"])
        except Exception as e:  # noqa: BLE001 - the reference's message (:213)
            raise RuntimeError(f"Failed to read video info: {e}") from e
        self.hwaccel = self._detect_hwaccel()
        self._dispatch = None
        self._lock = threading.Lock()

    def _detect_hwaccel(self) -> Optional[str]:
        import subprocess

        try:
            out = subprocess.run([self.ffmpeg, "-hwaccels"], capture_output=True, text=True, check=False).stdout
        except OSError:
            return None
        return "cuda" if "cuda" in out else "nvdec" if "nvdec" in out else None

    def _open(self) -> _PipeCapture:
        return _PipeCapture(self.ffmpeg, self.path, self.width, self.height, self.hwaccel)

    def __len__(self) -> int:
        return self.n

    def reader(self, long_ranges: bool = False):
        if not long_ranges:
            with self._lock:
                if self._dispatch is None:
                    self._dispatch = _SequentialDecoder(self.path, self.lookahead, self._open)
            return _SharedReader(self._dispatch)
        return _VideoReader(self.path, "grab", self._open)

    def close(self) -> None:
        if self._dispatch is not None:
            self._dispatch.stop()
            self._dispatch = None


class FfmpegPipeSink:
    """libx264 through the `ffmpeg` binary, fed raw BGR frames on stdin -- the reference's encoder (:514-532): `-crf`, `-preset`,
    yuv420p, +faststart. Frames must arrive in order (the reassembler guarantees it)."""

    def __init__(self, path: str, fps: float, crf: int = 15, preset: str = "slow", ffmpeg_bin: Optional[str] = None):
        found = find_ffmpeg()
        self.exe = ffmpeg_bin or (found[0] if found else None)
        if not self.exe:
            raise OSError("ffmpeg not found")
        self.path, self.fps, self.crf, self.preset = str(path), fps, int(crf), str(preset)
        self.proc = None

    def write(self, index: int, frame: np.ndarray) -> None:
        import subprocess

        if self.proc is None:
            h, w = frame.shape[:2]
            cmd = [self.exe, "-y", "-loglevel", "error", "-f", "rawvideo", "-vcodec", "rawvideo", "-s", f"{w}x{h}", "-pix_fmt", "bgr24",
                   "-r", str(self.fps), "-i", "-", "-an", "-vcodec", "libx264", "-crf", str(self.crf), "-preset", self.preset,
                   "-pix_fmt", "yuv420p", "-movflags", "+faststart", self.path]
            self.proc = subprocess.Popen(cmd, stdin=subprocess.PIPE, stderr=subprocess.DEVNULL)
        try:
            self.proc.stdin.write(np.ascontiguousarray(frame).data)
        except BrokenPipeError as e:
            raise OSError(f"ffmpeg stopped reading while encoding {self.path} (exit {self.proc.poll()})") from e

    def close(self) -> None:
        if self.proc is not None:
            self.proc.stdin.close()
            rc = self.proc.wait()
            self.proc = None
            if rc != 0:
                raise OSError(f"ffmpeg exited with {rc} while encoding {self.path}")


def open_video_source(path: str):
    """The reference's ffmpeg pipe when the binaries are on PATH, OpenCV otherwise (VR_IO=cv2 forces OpenCV)."""
    return FfmpegPipeSource(path) if find_ffmpeg() else VideoFileSource(path)


def open_video_sink(path: str, fps: float, crf: int = 15, preset: str = "slow"):
    return FfmpegPipeSink(path, fps, crf, preset) if find_ffmpeg() else VideoFileSink(path, fps)


class NullSink:
    """Counts frames and keeps an order-sensitive checksum (benchmarks, tests)."""

    def __init__(self):
        self.count = 0
        self.checksum = 0
        self.order: List[int] = []

    def write(self, index: int, frame: np.ndarray) -> None:
        self.count += 1
        self.order.append(index)
        self.checksum = (self.checksum * 1000003 + int(frame[::97, ::89].astype(np.uint32).sum()) + index) % (1 << 61)

    def close(self) -> None:
        pass


class ListSink:
    def __init__(self):
        self.frames: List[np.ndarray] = []
        self.order: List[int] = []

    def write(self, index: int, frame: np.ndarray) -> None:
        self.order.append(index)
        self.frames.append(frame.copy())

    def close(self) -> None:
        pass


class VideoFileSink:
    """cv2.VideoWriter (stands in for the reference's libx264 encoder pipe, :514-532)."""

    def __init__(self, path: str, fps: float, fourcc: str = "mp4v"):
        self.path, self.fps, self.fourcc = str(path), fps, fourcc
        self.writer = None

    def write(self, index: int, frame: np.ndarray) -> None:
        import cv2

        if self.writer is None:
            self.writer = cv2.VideoWriter(self.path, cv2.VideoWriter_fourcc(*self.fourcc), self.fps,
                                          (frame.shape[1], frame.shape[0]))
            if not self.writer.isOpened():
                raise OSError(f"cannot open {self.path} for writing")
        self.writer.write(frame)

    def close(self) -> None:
        if self.writer is not None:
            self.writer.release()
            self.writer = None


# ---------------------------------------------------------------------------------------------------------------
# chunk plan
# ---------------------------------------------------------------------------------------------------------------
def plan_chunks(total: int, n_workers: int, chunk: Optional[int] = None) -> List[tuple]:
    """[(start, end, worker)] covering [0, total): chunk c -> worker c % n_workers. chunk=None: one contiguous range
    per worker (sizes differ by at most one frame, like sharder.shard_range)."""
    if total < 0 or n_workers <= 0:
        raise ValueError("bad total / worker count")
    if chunk is None:
        from .sharder import shard_range

        out = []
        for w in range(n_workers):
            s, e = shard_range(total, w, n_workers)
            if e > s:
                out.append((s, e, w))
        return out
    if chunk <= 0:
        raise ValueError("chunk must be positive")
    return [(s, min(s + chunk, total), c % n_workers) for c, s in enumerate(range(0, total, chunk))]


# ---------------------------------------------------------------------------------------------------------------
# frame buffers
# ---------------------------------------------------------------------------------------------------------------
class BufferPool:
    """At most `count` page-locked output frames, allocated on first use. Restorers render straight into them
    (FrameRestorer.process_stream(out_pool=...)), the reassembler hands them to the sink and gives them back: no
    per-frame copy, and the host memory of the whole pipeline is bounded by count * frame bytes."""

    def __init__(self, count: int, alloc: Callable, min_count: int = 0, max_bytes: Optional[int] = None):
        """count: upper bound on buffers; max_bytes: soft cap on page-locked memory -- once `min_count` buffers exist (what the
        reorder ring needs to make progress) no further buffer is allocated beyond it; callers wait for a release instead.
        (the pool used to pin capacity + 5 G frames whatever their size, ~17 GB for 8 GPUs at 4K output.)"""
        self.count, self._alloc = count, alloc
        self.min_count, self.max_bytes = min_count, max_bytes
        self._cv = threading.Condition()
        self._free: List[np.ndarray] = []
        self.allocated = 0
        self._abort = False

    def get(self, shape) -> np.ndarray:
        shape = tuple(shape)
        with self._cv:
            while True:
                if self._abort:
                    raise RuntimeError("pipeline aborted")
                for i, b in enumerate(self._free):
                    if b.shape == shape:
                        return self._free.pop(i)
                if self.allocated < self.count and self._within_bytes(shape):
                    self.allocated += 1
                    break
                if self._free:  # another frame size (a new clip): drop a stale buffer and allocate
                    self._free.pop()
                    break
                self._cv.wait(0.5)
        return self._alloc(shape)

    def _within_bytes(self, shape) -> bool:
        if self.max_bytes is None or self.allocated < max(self.min_count, 1):
            return True
        return (self.allocated + 1) * int(np.prod(shape)) <= self.max_bytes

    def reserve(self, shape, n: int) -> None:
        """Allocate up to `n` more buffers now: page-locked allocations synchronise the device(s), so a pool that grows
        while frames are in flight stalls every GPU for ~10 ms per buffer."""
        shape = tuple(shape)
        for _ in range(n):
            with self._cv:
                if self.allocated >= self.count or not self._within_bytes(shape):
                    return
                self.allocated += 1
            buf = self._alloc(shape)
            self.release(buf)

    def release(self, buf: np.ndarray) -> None:
        with self._cv:
            self._free.append(buf)
            self._cv.notify()

    def abort(self) -> None:
        with self._cv:
            self._abort = True
            self._cv.notify_all()

    def close(self) -> None:
        """Drop the buffers now. Page-locked memory is returned by a finalizer (cudaFreeHost synchronises the device):
        left to the garbage collector it would fire in the middle of somebody's next run."""
        import gc

        with self._cv:
            self._free.clear()
        gc.collect()


# ---------------------------------------------------------------------------------------------------------------
# ordered reassembly
# ---------------------------------------------------------------------------------------------------------------
class OrderedReassembler:
    """Bounded reorder ring in front of a sequential sink. `put` copies the frame (restorers hand out views of pinned
    buffers that are reused) and blocks while the index is more than `capacity` frames ahead of the writer."""

    def __init__(self, sink, total: int, capacity: int):
        if capacity < 1:
            raise ValueError("capacity must be >= 1")
        self.sink, self.total, self.capacity = sink, total, capacity
        self._cv = threading.Condition()
        self._slots: dict = {}
        self._free: List[np.ndarray] = []
        self._next = 0
        self._error: Optional[BaseException] = None
        self.max_held = 0
        self._thread = threading.Thread(target=self._writer, name="vr-writer", daemon=True)
        self._thread.start()

    def put(self, index: int, frame: np.ndarray, release: Optional[Callable] = None) -> None:
        """Without `release` the frame is copied (it may be a view of a buffer the producer reuses); with it the
        reassembler takes the buffer over and calls release(frame) once the sink has written it."""
        with self._cv:
            while index >= self._next + self.capacity and index < self.total and self._error is None:
                self._cv.wait(0.5)
            if self._error is not None:
                raise RuntimeError("reassembly aborted") from self._error
            if index >= self.total:  # the source ended early (truncate): frames past the end are dropped
                if release is not None:
                    release(frame)
                return
            buf = None
            if release is None and self._free and self._free[-1].shape == frame.shape:
                buf = self._free.pop()
        if release is None:
            if buf is None:
                buf = np.empty_like(frame)
            np.copyto(buf, frame)
        else:
            buf = frame
        with self._cv:
            self._slots[index] = (buf, release)
            self.max_held = max(self.max_held, len(self._slots))
            self._cv.notify_all()

    def abort(self, exc: BaseException) -> None:
        with self._cv:
            if self._error is None:
                self._error = exc
            self._cv.notify_all()

    def truncate(self, total: int) -> None:
        """The source delivered fewer frames than it announced (container frame counts are estimates): stop after
        frame total - 1 instead of waiting for frames that will never come."""
        with self._cv:
            if total < self.total:
                self.total = total
                for i in [k for k in self._slots if k >= total]:
                    buf, release = self._slots.pop(i)
                    if release is not None:
                        release(buf)
            self._cv.notify_all()

    def _writer(self) -> None:
        try:
            while True:
                with self._cv:
                    while self._next not in self._slots and self._next < self.total and self._error is None:
                        self._cv.wait(0.5)
                    if self._error is not None or self._next >= self.total:
                        return
                    buf, release = self._slots.pop(self._next)
                    idx = self._next
                self.sink.write(idx, buf)
                if release is not None:
                    release(buf)
                with self._cv:
                    self._next += 1
                    if release is None and len(self._free) < 4:
                        self._free.append(buf)
                    self._cv.notify_all()
        except BaseException as e:  # noqa: BLE001 - surfaced to the workers and to finish()
            self.abort(e)

    def finish(self) -> None:
        self._thread.join()
        if self._error is not None:
            raise RuntimeError("pipeline failed") from self._error


# ---------------------------------------------------------------------------------------------------------------
# boundary frames
# ---------------------------------------------------------------------------------------------------------------
class _DeviceBoundary:
    """A boundary frame that already sits in device memory of the consuming restorer's GPU (FrameRestorer.boundary_send)."""

    def __init__(self, ptr: int):
        self.ptr = ptr


class _Boundaries:
    """u_last of chunk c, published by its worker for the worker of chunk c + 1 (one frame per chunk boundary)."""

    def __init__(self):
        self._cv = threading.Condition()
        self._frames: dict = {}
        self._error = False

    def publish(self, chunk_index: int, frame: np.ndarray) -> None:
        with self._cv:
            self._frames[chunk_index] = frame
            self._cv.notify_all()

    def take(self, chunk_index: int) -> np.ndarray:
        with self._cv:
            while chunk_index not in self._frames and not self._error:
                self._cv.wait(0.5)
            if self._error:
                raise RuntimeError("pipeline aborted while waiting for a boundary frame")
            return self._frames.pop(chunk_index)

    def abort(self) -> None:
        with self._cv:
            self._error = True
            self._cv.notify_all()


@dataclass
class PipelineStats:
    frames: int = 0
    seconds: float = 0.0
    boundary_frames: int = 0
    chunks: int = 0
    max_held: int = 0
    setup_seconds: float = 0.0

    @property
    def fps(self) -> float:
        return self.frames / self.seconds if self.seconds > 0 else 0.0


def _default_temporal_blend(gpu_id: int):
    from .restorer import temporal_blend

    return lambda cur, prev, alpha, tau: temporal_blend(cur, prev, alpha, tau, device=gpu_id)


def run_pipeline(source, sink, make_restorer: Callable[[int], object], gpu_ids: Sequence[int], opts,
                 chunk: Optional[int] = None, capacity: Optional[int] = None,
                 temporal_blend: Optional[Callable[[int], Callable]] = None, warmup: bool = True) -> PipelineStats:
    """Restore every frame of `source` into `sink` (in order) on the GPUs `gpu_ids` (one thread + one restorer each;
    an id may repeat to run two restorers on one GPU). `make_restorer(gpu_id)` builds a FrameRestorer-like object
    (process_stream, temporal_reset, temporal_get_prev, scale, close); `opts` is a FrameOpts.

    chunk: frames per contiguous range; None = ceil(F / G) (one range per GPU, `sharder.py`'s split), which needs a
    reorder ring of the whole video -- a sequential sink wants a small chunk (the CLI uses 16).
    capacity: reorder-ring frames; with the temporal stage and more than one chunk the minimum is G * C (default
    G * C + G), otherwise any value >= 1 works (default min(G * C, 64)).
    warmup: run one frame through every restorer before the clock starts (counted as set-up)."""
    total = len(source)
    gpu_ids = list(gpu_ids)
    if not gpu_ids:
        raise ValueError("no GPUs given")
    G = len(gpu_ids)
    plan = plan_chunks(total, G, chunk)
    C_max = max((e - s for s, e, _ in plan), default=1)
    use_temporal = bool(getattr(opts, "temporal", False))
    # a deferred head frame is written after the rest of its chunk: the ring must then hold one chunk per GPU
    need = G * C_max if (use_temporal and len(plan) > 1) else 1
    if capacity is None:
        capacity = need + G if need > 1 else max(min(G * C_max, 64), 4)
    if capacity < need:
        raise ValueError(f"capacity {capacity} < {need}: a chunk's head frame is written last, so the ring must hold "
                         f"one chunk per GPU")
    reasm = OrderedReassembler(sink, total, capacity)
    bounds = _Boundaries()
    make_blend = temporal_blend or _default_temporal_blend
    peer = temporal_blend is None  # a caller-supplied blend (tests, stubs) keeps the host path
    stats = PipelineStats(chunks=len(plan))
    errors: List[BaseException] = []
    lock = threading.Lock()

    def worker(slot: int) -> None:
        gpu = gpu_ids[slot]
        restorer = None
        try:
            restorer = make_restorer(gpu)
            restorers[slot] = restorer
            blend = make_blend(gpu)
            try:
                reader = source.reader(long_ranges=chunk is None and G > 1)
            except TypeError:  # sources without the keyword (synthetic / array / user-supplied)
                reader = source.reader()
            zero_copy = bool(getattr(restorer, "zero_copy_stream", False))
            if zero_copy:
                with lock:
                    if pool[0] is None:
                        # ring + per worker: two frames in flight, one being handed over, the deferred head, a boundary frame
                        # soft cap on pinned memory (VR_PINNED_MB, default 8192): never below what the ring needs to progress
                        import os
                        cap_mb = int(os.environ.get("VR_PINNED_MB", "8192"))
                        pool[0] = BufferPool(capacity + 5 * G, restorer.alloc_host, min_count=need + 2 * G + 1,
                                             max_bytes=cap_mb << 20)
            kw = {"out_pool": pool[0]} if zero_copy else {}
            give_back = pool[0].release if zero_copy else None
            if warmup and total > 0:
                # first-use costs (activation buffers, tensor maps, the pinned frame rings) belong to the set-up
                for out in restorer.process_stream(reader.read_range(plan[0][0], plan[0][0] + 1), opts, **kw):
                    if give_back:
                        give_back(out)
                        # this worker's share of the pool, up front (with deferred head frames the ring does fill up)
                        want = (capacity + G - 1) // G + 5 if need > 1 else 8
                        pool[0].reserve(out.shape, want - 1)
            ready.wait()  # the clock starts when every GPU has its weights (model set-up is not frame throughput)
            if slot == 0:
                t_start[0] = time.perf_counter()
            for ci, (s, e, w) in enumerate(plan):
                if w != slot:
                    continue
                defer_head = use_temporal and ci > 0
                head = None
                if use_temporal:
                    restorer.temporal_reset()
                out_shape = None
                got = 0
                for k, out in enumerate(restorer.process_stream(reader.read_range(s, e), opts, **kw)):
                    out_shape = out.shape
                    got = k + 1
                    if k == 0 and defer_head:
                        # u_first: passed through by the reset temporal stage; kept until the boundary frame is here
                        head = out if zero_copy else out.copy()
                    else:
                        reasm.put(s + k, out, give_back)
                if got < e - s:
                    reasm.truncate(s + got)
                if use_temporal and ci + 1 < len(plan) and out_shape is None:
                    bounds.publish(ci, None)  # nothing decoded: the next chunk (past the end) must not wait
                if use_temporal and ci + 1 < len(plan) and out_shape is not None:
                    nxt = restorers[plan[ci + 1][2]]
                    if peer and getattr(nxt, "peer_boundary", False):
                        # ONE cudaMemcpyPeerAsync: this GPU's last un-blended frame -> a device buffer on the next chunk's GPU
                        b = _DeviceBoundary(restorer.boundary_send(nxt, out_shape[0], out_shape[1]))
                    elif zero_copy:
                        b = pool[0].get(out_shape)
                        restorer.temporal_get_prev(out_shape[0], out_shape[1], out=b)
                    else:
                        b = restorer.temporal_get_prev(out_shape[0], out_shape[1])
                    bounds.publish(ci, b)
                if defer_head:
                    prev = bounds.take(ci - 1)
                    if isinstance(prev, _DeviceBoundary):
                        if head is None:
                            restorer.boundary_finish(prev.ptr, None, None)   # nothing decoded for this chunk: recycle only
                        else:
                            with lock:
                                stats.boundary_frames += 1
                            dst = pool[0].get(head.shape) if zero_copy else np.empty_like(head)
                            restorer.boundary_finish(prev.ptr, head, dst, opts.temporal_alpha, opts.temporal_tau)
                            reasm.put(s, dst, give_back)
                            if zero_copy:
                                give_back(head)
                    elif head is None:
                        if prev is not None and zero_copy:
                            give_back(prev)
                    elif prev is None:
                        reasm.put(s, head, give_back)
                    else:
                        with lock:
                            stats.boundary_frames += 1
                        reasm.put(s, blend(head, prev, opts.temporal_alpha, opts.temporal_tau))
                        if zero_copy:
                            give_back(head)
                            give_back(prev)
        except BaseException as exc:  # noqa: BLE001 - propagate to every thread, re-raised by the caller
            with lock:
                errors.append(exc)
            ready.abort()
            bounds.abort()
            if pool[0] is not None:
                pool[0].abort()
            reasm.abort(exc)
        finally:
            if restorer is not None and hasattr(restorer, "close"):
                restorer.close()

    t0 = time.perf_counter()
    t_start = [t0]
    pool: List[Optional[BufferPool]] = [None]
    restorers: List[object] = [None] * G   # filled by the workers before the start barrier
    ready = threading.Barrier(G)
    threads = [threading.Thread(target=worker, args=(i,), name=f"vr-gpu{gpu_ids[i]}", daemon=True) for i in range(G)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:  # the first failure wins; BrokenBarrierError / "aborted" from the other threads are its echoes
        reasm.abort(errors[0])
        reasm._thread.join()
        sink.close()
        first = next((e for e in errors if not isinstance(e, threading.BrokenBarrierError)), errors[0])
        raise first
    try:
        reasm.finish()
    finally:
        sink.close()
    t_end = time.perf_counter()
    stats.seconds = t_end - t_start[0]
    stats.setup_seconds = t_start[0] - t0
    stats.frames = reasm.total
    stats.max_held = reasm.max_held
    if pool[0] is not None:
        reasm._free.clear()
        pool[0].close()
    if hasattr(source, "close"):
        source.close()
    return stats
