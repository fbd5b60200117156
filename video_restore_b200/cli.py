"""Command-line surface of the reference's video_upscaler.py (argparse at video_upscaler.py:649-682, presets at
:687-701, config at :704-718), driving the B200 hot path. Only the flag surface and the per-frame stage are
reproduced; the progress bar (video_upscaler.py:569-602) is out of scope. Video I/O is the
reference's ffmpeg rawvideo pipes (pipeline.FfmpegPipeSource / FfmpegPipeSink: `-hwaccel` decode :220-262, libx264
`-crf` / `-preset` encode :514-532) and its audio mux (:604-627, `copy_audio`) when the `ffmpeg` and `ffprobe` binaries
are on PATH, otherwise OpenCV VideoCapture / VideoWriter. Frames flow through the in-process multi-GPU pipeline of
pipeline.py (one thread + restorer per `--gpus` id, contiguous frame chunks, ordered bounded reassembly: N2),
`--batch` walks a directory (N4), `--synthetic N` runs N generated 720p frames without any video I/O.

Flags added on top of the reference's parser are the README-only ones the north star names:
  --model RealESRGAN_x2plus (README.md:158), --denoise S, --sharpen A (README.md:140-141),
  --no-seamless / --no-temporal / --no-color-enhance (README.md:147-149).
"""
from __future__ import annotations

import argparse
import shutil
import sys
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional

MODELS = ["RealESRGAN_x4plus", "RealESRGAN_x4_v3", "RealESRGAN_x4plus_anime_6B", "RealESRGAN_x2plus"]


@dataclass
class OptimizedConfig:
    """Field-for-field mirror of the reference dataclass (video_upscaler.py:112-135) plus the README-only stage."""
    model_name: str = "RealESRGAN_x4plus"
    scale: int = 4
    gpu_ids: List[int] = field(default_factory=list)
    tile_size: int = 512
    tile_overlap: int = 32
    use_fp16: bool = True
    enhanced_mode: bool = False
    light_denoise: bool = False
    output_format: str = "mp4"
    crf: int = 15
    preset: str = "slow"
    audio_copy: bool = True
    prefetch_frames: int = 32
    # README-only enhancement stage
    seamless: bool = False
    temporal: bool = False
    color_enhance: bool = False
    denoise_strength: Optional[float] = None
    sharpen: float = 0.0

    @property
    def tile_pad(self) -> int:
        return self.tile_overlap if self.enhanced_mode else 10  # video_upscaler.py:326


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="B200-native AI video upscaler (video-restore hot path)")
    p.add_argument("input", help="Input video or directory")
    p.add_argument("output", help="Output video or directory")
    p.add_argument("--model", default="RealESRGAN_x4plus", choices=MODELS)
    p.add_argument("--gpus", type=int, nargs="+", default=None, help="GPU IDs to use (default: all)")
    p.add_argument("--quality", choices=["fast", "balanced", "max"], default="balanced")
    p.add_argument("--enhanced", action="store_true", help="Enable artifact reduction features")
    p.add_argument("--tile-size", type=int, default=None)
    p.add_argument("--tile-overlap", type=int, default=None)
    p.add_argument("--crf", type=int, default=None)
    p.add_argument("--preset", default=None, choices=["ultrafast", "fast", "medium", "slow", "veryslow"])
    p.add_argument("--no-audio", action="store_true")
    p.add_argument("--batch", action="store_true")
    p.add_argument("--denoise", type=float, default=None, help="bilateral strength (0.15 == the reference's 5/25/25)")
    p.add_argument("--sharpen", type=float, default=None, help="unsharp-mask amount")
    p.add_argument("--no-seamless", action="store_true")
    p.add_argument("--no-temporal", action="store_true")
    p.add_argument("--no-color-enhance", action="store_true")
    p.add_argument("--synthetic", type=int, default=0, help="process N synthetic 720p frames instead of a video")
    p.add_argument("--random-weights", action="store_true",
                   help="run with random-init weights when models/<model>.pth is absent (benchmarks / tests only: the output "
                        "is not a restored video)")
    p.add_argument("--procs", action="store_true",
                   help="with several --gpus: one PROCESS per GPU over contiguous frame ranges (scales like bench.py; a video "
                        "file is written as one segment per GPU) instead of one thread per GPU in this process")
    return p


def config_from_args(args) -> OptimizedConfig:
    """Quality presets exactly as video_upscaler.py:687-701 (including its `x or default` idiom)."""
    if args.quality == "max":
        crf = args.crf or 12
        preset = args.preset or "veryslow"
        tile_size = args.tile_size or (512 if args.enhanced else 1536)
        tile_overlap = args.tile_overlap or (64 if args.enhanced else 32)
    elif args.quality == "fast":
        crf = args.crf or 18
        preset = args.preset or "fast"
        tile_size = args.tile_size or 1024
        tile_overlap = args.tile_overlap or 16
    else:
        crf = args.crf or 15
        preset = args.preset or "slow"
        tile_size = args.tile_size or (512 if args.enhanced else 1024)
        tile_overlap = args.tile_overlap or (32 if args.enhanced else 16)
    cfg = OptimizedConfig(
        model_name=args.model, gpu_ids=args.gpus or [], tile_size=tile_size, tile_overlap=tile_overlap, crf=crf,
        preset=preset, audio_copy=not args.no_audio, enhanced_mode=args.enhanced, light_denoise=args.enhanced,
        use_fp16=True,
        seamless=args.enhanced and not args.no_seamless, temporal=args.enhanced and not args.no_temporal,
        color_enhance=args.enhanced and not args.no_color_enhance, denoise_strength=args.denoise,
        sharpen=args.sharpen if args.sharpen is not None else (0.1 if args.enhanced else 0.0))
    cfg.scale = 2 if args.model == "RealESRGAN_x2plus" else 4  # the reference hard-codes 4 (:718); x2plus is 2x
    return cfg


def frame_opts_from_config(cfg: OptimizedConfig):
    from .restorer import FrameOpts

    denoise = (cfg.enhanced_mode and cfg.light_denoise) or cfg.denoise_strength is not None
    sigma = 25.0
    if cfg.denoise_strength is not None:
        sigma = min(max(25.0 * cfg.denoise_strength / 0.15, 1.0), 150.0)
    return FrameOpts(denoise=denoise, denoise_d=5, denoise_sigma_color=sigma, denoise_sigma_space=sigma,
                     sharpen=cfg.sharpen, clahe=cfg.color_enhance, temporal=cfg.temporal)


class MissingWeights(FileNotFoundError):
    pass


def load_weights(cfg: OptimizedConfig, allow_random: bool = False):
    """state_dict of `cfg.model_name` from the reference's cache location models/<name>.pth (video_upscaler.py:350-353).
    The reference downloads a missing checkpoint (:342-367) or fails; there is no network here, so a missing file is an ERROR
    unless random weights were asked for explicitly (--random-weights, implied by --synthetic)."""
    path = Path("models") / f"{cfg.model_name}.pth"
    if path.exists():
        import torch
        net = torch.load(path, map_location="cpu")
        return net.get("params_ema", net.get("params", net))
    if not allow_random:
        raise MissingWeights(f"{path} not found (the reference would download it, video_upscaler.py:342-367; no network here). "
                             f"Put the checkpoint there, or pass --random-weights to run with random-init weights "
                             f"(benchmarks only: the output is NOT a restored video)")
    from .synth import random_state_dict

    print(f"[video-restore] {path} not found: random-init weights as requested (output is not a restored video)")
    return random_state_dict(cfg.model_name, seed=0)


def make_restorer(cfg: OptimizedConfig, gpu_id: int, state_dict=None, allow_random: bool = False):
    from .restorer import FrameRestorer

    if state_dict is None:
        state_dict = load_weights(cfg, allow_random)
    return FrameRestorer(cfg.model_name, state_dict, tile=cfg.tile_size, tile_pad=cfg.tile_pad,
                         blend="gaussian" if cfg.seamless else "crop", gpu_id=gpu_id)


VIDEO_EXTS = (".mp4", ".avi", ".mov", ".mkv", ".webm")  # the reference's batch filter, video_upscaler.py:732


def _run_one(cfg: OptimizedConfig, opts, source, sink, chunk: int, state_dict):
    """One clip through the multi-GPU pipeline (pipeline.py): one host thread + restorer per GPU in cfg.gpu_ids,
    contiguous frame chunks, one boundary frame per chunk for the temporal stage, ordered bounded reassembly."""
    from .pipeline import run_pipeline

    return run_pipeline(source, sink, lambda gpu_id: make_restorer(cfg, gpu_id, state_dict), cfg.gpu_ids, opts, chunk=chunk)


def copy_audio(input_path: str, output_path: str, ffmpeg_bin: str | None = None) -> bool:
    """Mux the source's audio track into the written video without re-encoding either stream -- what the reference's
    `_copy_audio` does through ffmpeg-python (video_upscaler.py:604-627): video stream of `output_path` + audio stream of
    `input_path` -> temp file -> replaces `output_path`. Like the reference, a clip without an audio track (or any ffmpeg
    failure) leaves the video as written and removes the temp file. Returns True when the mux happened. Needs the
    `ffmpeg` binary on PATH; returns False (video kept) when it is absent."""
    import os
    import shutil
    import subprocess

    exe = ffmpeg_bin or shutil.which("ffmpeg")
    if not exe:
        return False
    root, ext = os.path.splitext(output_path)
    temp_path = output_path + ".temp" + (ext or ".mp4")  # the reference's name: output + '.temp.mp4'
    cmd = [exe, "-y", "-loglevel", "error", "-i", output_path, "-i", input_path, "-map", "0:v", "-map", "1:a",
           "-c:v", "copy", "-c:a", "copy", temp_path]
    try:
        done = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False)
        if done.returncode == 0 and os.path.exists(temp_path) and os.path.getsize(temp_path) > 0:
            os.replace(temp_path, output_path)
            return True
    except OSError:
        pass
    if os.path.exists(temp_path):
        os.remove(temp_path)
    return False


def join_segments(segments, output_path: str, ffmpeg_bin: str | None = None) -> bool:
    """Concatenate the per-rank video segments of `--procs` (consecutive frame ranges, same codec and size) into one file
    with ffmpeg's concat demuxer, streams copied. Returns False (segments kept) without the `ffmpeg` binary or on failure."""
    import os
    import subprocess
    import tempfile

    exe = ffmpeg_bin or shutil.which("ffmpeg")
    if not exe or not segments:
        return False
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        for seg in segments:
            f.write("file '%s'\n" % os.path.abspath(seg).replace("'", "'\\''"))
        listing = f.name
    try:
        done = subprocess.run([exe, "-y", "-loglevel", "error", "-f", "concat", "-safe", "0", "-i", listing, "-c", "copy",
                               output_path], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False)
        ok = done.returncode == 0 and os.path.exists(output_path) and os.path.getsize(output_path) > 0
    except OSError:
        ok = False
    os.remove(listing)
    if ok:
        for seg in segments:
            if os.path.abspath(seg) != os.path.abspath(output_path) and os.path.exists(seg):
                os.remove(seg)
    return ok


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    cfg = config_from_args(args)
    opts = frame_opts_from_config(cfg)
    import torch

    from .pipeline import NullSink, SyntheticSource, find_ffmpeg, open_video_sink, open_video_source

    if not cfg.gpu_ids:
        cfg.gpu_ids = list(range(torch.cuda.device_count()))
    if not cfg.gpu_ids:
        print("Error: No CUDA GPUs available")  # same failure as video_upscaler.py:140-141
        return 1
    try:
        state_dict = load_weights(cfg, allow_random=args.random_weights or args.synthetic > 0)
    except MissingWeights as e:
        print(f"Error: {e}")
        return 1
    if args.procs and len(cfg.gpu_ids) > 1:
        from .multiproc import run_processes

        return run_processes(args, cfg, opts)
    # frames per contiguous range: bounds the reorder ring in front of the sequential encoder (G * 16 frames); one GPU
    # walks the clip as a single range
    chunk = 16 if len(cfg.gpu_ids) > 1 else None
    if args.synthetic > 0:
        st = _run_one(cfg, opts, SyntheticSource(720, 1280, args.synthetic, seed=1, distinct=8), NullSink(), chunk, state_dict)
        print(f"processed {st.frames} frames in {st.seconds:.2f} s ({st.fps:.2f} fps) on {len(cfg.gpu_ids)} GPU(s); "
              f"{st.boundary_frames} boundary frames exchanged; model set-up {st.setup_seconds:.1f} s")
        return 0
    # video I/O: the reference's ffmpeg rawvideo pipes (libx264, --crf / --preset honoured, audio copied) when the binaries are on
    # PATH; otherwise OpenCV -- and then what is NOT done with the reference's flags is said once, loudly
    if find_ffmpeg():
        print(f"[video-restore] video I/O through ffmpeg pipes: libx264 -crf {cfg.crf} -preset {cfg.preset}"
              + ("" if args.no_audio else ", audio copied from the source"))
    else:
        ignored = [f for f, v in (("--crf", args.crf), ("--preset", args.preset)) if v is not None]
        print("[video-restore] note: no ffmpeg / ffprobe on PATH -- video is read and written with OpenCV ('mp4v'), not libx264"
              + (f" -- {', '.join(ignored)} ignored" if ignored else "")
              + ("" if args.no_audio else "; audio is NOT copied (video_upscaler.py:604-627)"))
    if args.enhanced:
        print("[video-restore] note: --enhanced also enables the README's seamless blend / temporal / CLAHE / unsharp stage, which "
              "the reference's code does not implement (--no-seamless --no-temporal --no-color-enhance --sharpen 0 turn it off)")
    jobs = []
    if args.batch:  # directory mode, video_upscaler.py:723-746
        in_dir, out_dir = Path(args.input), Path(args.output)
        if not in_dir.is_dir():
            print(f"Error: {in_dir} is not a directory")
            return 1
        out_dir.mkdir(parents=True, exist_ok=True)
        for f in sorted(in_dir.iterdir()):
            if f.suffix.lower() in VIDEO_EXTS:
                jobs.append((f, out_dir / f"{f.stem}_upscaled{f.suffix}"))  # source suffix kept, as :744
        if not jobs:
            print(f"No videos found in {in_dir}")
            return 1
        print(f"\nBatch processing {len(jobs)} videos\n")
    else:
        jobs.append((Path(args.input), Path(args.output)))
    rc = 0
    for src_path, dst_path in jobs:
        # like process_video (:369-428): a failing clip is reported and the batch goes on
        try:
            source = open_video_source(str(src_path))
            st = _run_one(cfg, opts, source, open_video_sink(str(dst_path), source.fps, cfg.crf, cfg.preset), chunk, state_dict)
        except KeyboardInterrupt:
            print("\n\nProcessing interrupted")
            return 1
        except Exception as e:  # noqa: BLE001
            print(f"Error: {src_path.name}: {e}")
            rc = 1
            continue
        if st.frames == 0:
            print(f"Error: {src_path.name}: no frames decoded")
            rc = 1
            continue
        muxed = cfg.audio_copy and copy_audio(str(src_path), str(dst_path))  # :421-423
        print(f"{src_path.name}: processed {st.frames} frames in {st.seconds:.2f} s ({st.fps:.2f} fps) on "
              f"{len(cfg.gpu_ids)} GPU(s)" + ("; audio copied" if muxed else ""))
    return rc


if __name__ == "__main__":
    sys.exit(main())
