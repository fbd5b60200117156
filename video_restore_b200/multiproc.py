"""`--gpus a b c ... --procs`: one PROCESS per GPU over contiguous frame ranges -- the CLI path that scales like bench.py.

Replaces the reference's thread-per-GPU workers on one shared queue (video_upscaler.py:388-394, :453-488). The in-process
pipeline (pipeline.py, one thread per GPU) shares one GIL and one allocator between all GPUs and measured 1.63x on two GPUs
in round 1; one process per GPU measured 1.97x. Here every rank

  * owns frames [start, end) = sharder.shard_range(total, rank, world) and its own decoder (exact forward skipping, no seeks),
  * restores them through FrameRestorer.process_stream (H2D / compute / D2H overlapped),
  * exchanges ONE boundary frame with its neighbours, device to device over NVLink (sharder.FrameRangeSharder, nccl:
    grouped isend/irecv, all boundaries concurrently), when the temporal stage is on,
  * writes its own sink: a frame digest for `--synthetic`, or a video SEGMENT `<out stem>.part<rank><suffix>` -- the encoders
    run in parallel too, which matters more than the GPUs once frames come at > 100 per second. Segments are complete,
    independently playable files covering consecutive frame ranges; joining them is a stream copy
    (`ffmpeg -f concat -i list.txt -c copy out.mp4`, list printed at the end) that needs the ffmpeg binary, absent here.

torch.distributed (nccl) is plumbing only: rendezvous, one barrier before the clock starts, the boundary frame, and a max-
reduction of the elapsed time. Results are bit-identical to the single-GPU run (tests/test_pipeline.py, `-m gpu`, skipped with
fewer than two devices).
"""
from __future__ import annotations

import os
import socket
import time
import zlib
from pathlib import Path

import numpy as np


def frame_digest(frame: np.ndarray) -> int:
    """crc32 of every 8th row (cheap enough for > 100 frames/s in one thread; sensitive to any row it covers)."""
    return zlib.crc32(np.ascontiguousarray(frame[::8]).tobytes())


def combine_digests(per_frame: dict) -> str:
    """Order-defined digest of a clip from its per-frame digests {index: crc32}."""
    h = 0
    for i in sorted(per_frame):
        h = zlib.crc32(f"{i}:{per_frame[i]};".encode(), h)
    return f"{h:08x}"


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, job: dict, result_q) -> None:
    import torch
    import torch.distributed as dist

    from .cli import OptimizedConfig, make_restorer
    from .pipeline import FfmpegPipeSink, SyntheticSource, VideoFileSink, _PipeCapture, _VideoReader
    from .restorer import FrameOpts
    from .sharder import FrameRangeSharder

    try:
        gpu = job["gpu_ids"][rank]
        torch.cuda.set_device(gpu)
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", gpu))
        cfg = OptimizedConfig(**job["cfg"])
        opts = FrameOpts(**job["opts"])
        restorer = make_restorer(cfg, gpu, allow_random=job["allow_random"])
        total = job["total"]
        sh = FrameRangeSharder(rank, world, total)
        io = job.get("io")  # {"ffmpeg", "hwaccel"} when the parent found the binaries: the reference's rawvideo pipes, else OpenCV

        def open_reader():
            # every rank decodes its own range with a decoder of its own and exact forward skipping
            if io:
                return _VideoReader(job["input"], "grab", lambda: _PipeCapture(io["ffmpeg"], job["input"], job["width"],
                                                                                 job["height"], io.get("hwaccel")))
            return _VideoReader(job["input"], "grab")

        if job["synthetic"]:
            src = SyntheticSource(job["height"], job["width"], total, seed=1, distinct=8)
            reader = src.reader()
        else:
            reader = open_reader()
        sink = None
        if job["output"]:
            out = Path(job["output"])
            seg = out.with_name(f"{out.stem}.part{rank:02d}{out.suffix}")
            sink = (FfmpegPipeSink(str(seg), job["fps"], cfg.crf, cfg.preset, io["ffmpeg"]) if io else
                    VideoFileSink(str(seg), job["fps"]))
        digests = {}
        # set-up outside the clock: weights are resident, one frame has gone through (buffers, tensor maps, pinned rings),
        # the point-to-point connections are open
        first = next(iter(reader.read_range(sh.start, sh.start + 1)), None) if sh.end > sh.start else None
        if first is not None:
            for _ in restorer.process_stream(iter([first]), opts):
                pass
        sh.connect()
        it = iter(reader.read_range(sh.start, sh.end))
        nxt = [sh.start]

        def get_frame(i):
            if i == nxt[0]:  # the sequential walk of the range
                nxt[0] += 1
                return next(it)
            # random access (the boundary frame of the in-order protocol): a decoder of its own, exact forward skipping
            if job["synthetic"]:
                return next(iter(src.read_range(i, i + 1)))
            return next(iter(open_reader().read_range(i, i + 1)))

        def put_frame(i, o):
            digests[i] = frame_digest(o)
            if sink is not None:
                sink.write(i, o)

        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        # a file segment needs its frames in order, so its shard upscales the boundary frame first (one extra frame per
        # shard); a digest sink takes the head frame last (no redundant work)
        n = sh.run_stream(restorer, get_frame, put_frame, opts, defer_head=sink is None)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if sink is not None:
            sink.close()
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        result_q.put({"rank": rank, "frames": n, "seconds": float(t.item()), "own_seconds": dt, "digests": digests,
                      "exchange_ms": sh.exchange_ms, "segment": str(sink.path) if sink is not None else None})
        restorer.close()
        dist.barrier()
        dist.destroy_process_group()
    except BaseException as e:  # noqa: BLE001 - reported to the parent, which fails the job
        result_q.put({"rank": rank, "error": f"{type(e).__name__}: {e}"})
        raise


def run_job(gpu_ids, cfg, opts, total: int, synthetic: bool, input_path=None, output_path=None, fps: float = 30.0,
            height: int = 720, width: int = 1280, allow_random: bool = False, io: dict | None = None) -> dict:
    """Run one clip on `gpu_ids`, one process each. Returns {frames, seconds, fps, digest, per_frame, exchange_ms, segments}."""
    import dataclasses

    import torch.multiprocessing as mp

    world = len(gpu_ids)
    if total < world:
        raise ValueError(f"{total} frames for {world} processes: use fewer GPUs")
    job = {"gpu_ids": list(gpu_ids), "cfg": {**dataclasses.asdict(cfg), "gpu_ids": list(gpu_ids)},
           "opts": dataclasses.asdict(opts), "total": int(total), "synthetic": bool(synthetic), "input": input_path,
           "output": output_path, "fps": float(fps), "height": height, "width": width, "allow_random": allow_random, "io": io}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, job, q), daemon=False) for r in range(world)]
    for p in procs:
        p.start()
    results = []
    try:
        while len(results) < world:
            try:
                results.append(q.get(timeout=5.0))
            except Exception:  # noqa: BLE001 - queue.Empty: check that everybody is still alive
                if any(p.exitcode not in (None, 0) for p in procs):
                    break
            if results and "error" in results[-1]:
                break
    finally:
        failed = [r for r in results if "error" in r] or (len(results) < world)
        for p in procs:
            p.join(timeout=30.0 if not failed else 5.0)
            if p.is_alive():
                p.kill()
    errs = [r["error"] for r in results if "error" in r]
    if errs or len(results) < world:
        raise RuntimeError("multi-process run failed: " + ("; ".join(errs) if errs else "a worker exited early"))
    per_frame = {}
    for r in results:
        per_frame.update(r["digests"])
    seconds = max(r["seconds"] for r in results)
    frames = sum(r["frames"] for r in results)
    return {"frames": frames, "seconds": seconds, "fps": frames / seconds if seconds > 0 else 0.0,
            "digest": combine_digests(per_frame), "per_frame": per_frame,
            "exchange_ms": max(r["exchange_ms"] for r in results),
            "segments": [r["segment"] for r in sorted(results, key=lambda r: r["rank"]) if r["segment"]]}


def run_processes(args, cfg, opts) -> int:
    """CLI entry (cli.main with --procs)."""
    from .pipeline import FfmpegPipeSource, open_video_source

    allow_random = bool(args.random_weights or args.synthetic > 0)
    if args.synthetic > 0:
        res = run_job(cfg.gpu_ids, cfg, opts, args.synthetic, True, allow_random=allow_random)
        print(f"processed {res['frames']} frames in {res['seconds']:.2f} s ({res['fps']:.2f} fps) on {len(cfg.gpu_ids)} GPU(s), "
              f"one process per GPU; boundary exchange {res['exchange_ms']:.1f} ms; frames digest {res['digest']}")
        return 0
    jobs = []
    if args.batch:
        from .cli import VIDEO_EXTS

        in_dir, out_dir = Path(args.input), Path(args.output)
        if not in_dir.is_dir():
            print(f"Error: {in_dir} is not a directory")
            return 1
        out_dir.mkdir(parents=True, exist_ok=True)
        jobs = [(f, out_dir / f"{f.stem}_upscaled{f.suffix}") for f in sorted(in_dir.iterdir()) if f.suffix.lower() in VIDEO_EXTS]
        if not jobs:
            print(f"No videos found in {in_dir}")
            return 1
    else:
        jobs = [(Path(args.input), Path(args.output))]
    rc = 0
    for src_path, dst_path in jobs:
        try:
            src = open_video_source(str(src_path))  # counts the frames exactly, once, for every rank
            io = {"ffmpeg": src.ffmpeg, "hwaccel": src.hwaccel} if isinstance(src, FfmpegPipeSource) else None
            res = run_job(cfg.gpu_ids, cfg, opts, len(src), False, str(src_path), str(dst_path), src.fps, height=src.height,
                          width=src.width, allow_random=allow_random, io=io)
        except Exception as e:  # noqa: BLE001 - like the reference's process_video: report and go on
            print(f"Error: {src_path.name}: {e}")
            rc = 1
            continue
        print(f"{src_path.name}: processed {res['frames']} frames in {res['seconds']:.2f} s ({res['fps']:.2f} fps) on "
              f"{len(cfg.gpu_ids)} GPU(s), one process per GPU")
        from .cli import copy_audio, join_segments

        if join_segments(res["segments"], str(dst_path)):
            muxed = cfg.audio_copy and copy_audio(str(src_path), str(dst_path))
            print(f"  joined {len(res['segments'])} segments into {dst_path}" + ("; audio copied" if muxed else ""))
            continue
        print("segments (consecutive frame ranges; no ffmpeg binary on PATH -- join with `ffmpeg -f concat -safe 0 -i list.txt "
              f"-c copy {dst_path}`, list.txt = one `file '<segment>'` line each):")
        for s in res["segments"]:
            print(f"  {s}")
    return rc
