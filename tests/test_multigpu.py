"""Multi-device bit-identity on REAL hardware (`-m gpu`; every test SKIPS -- does not pass -- with fewer than two devices;
run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`, record under profiles/).

SURVEY.md section 4: "N-shard output identical (bit-exact) to 1-shard output, incl. the temporal boundary frame; run with
1/2/4/8 devices". Three paths over the same clip, all compared with one restorer walking the clip in order on device 0:
  * the in-process pipeline (pipeline.run_pipeline, one thread per GPU, boundary frame = one cudaMemcpyPeer),
  * one process per GPU (multiproc.run_job: FrameRangeSharder over nccl, device-resident grouped isend/irecv),
  * the CLI's `--procs` path end to end (frames digest printed by both the 1-GPU and the N-GPU run).
Replaces the reference's thread-per-GPU dispatch, video_upscaler.py:388-394, :453-488."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NAME, TILE, PAD = "RealESRGAN_x4_v3", 64, 10
H, W, N = 40, 56, 13


def _devices():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 CUDA devices, found {n}")
    return min(n, 8)


def _clip():
    from video_restore_b200.synth import synth_frame

    return [synth_frame(H, W, seed=9, index=i) for i in range(N)]


def _opts():
    from video_restore_b200.restorer import FrameOpts

    return FrameOpts(denoise=True, sharpen=0.3, clahe=True, temporal=True)


def _single(frames, sd):
    from video_restore_b200.restorer import FrameRestorer

    one = FrameRestorer(NAME, sd, tile=TILE, tile_pad=PAD, gpu_id=0)
    one.temporal_reset()
    want = [one.process_frame(f, _opts()) for f in frames]
    one.close()
    return want


def test_in_process_pipeline_on_distinct_devices(gpu_lib):
    from video_restore_b200.pipeline import ArraySource, ListSink, run_pipeline
    from video_restore_b200.restorer import FrameRestorer
    from video_restore_b200.synth import random_state_dict

    g = _devices()
    sd = random_state_dict(NAME, seed=0)
    frames = _clip()
    want = _single(frames, sd)
    for gpus, chunk in ((list(range(g)), 2), ([0, 1], None), (list(range(g)), None)):
        if len(gpus) > N:
            continue
        sink = ListSink()
        st = run_pipeline(ArraySource(frames), sink, lambda d: FrameRestorer(NAME, sd, tile=TILE, tile_pad=PAD, gpu_id=d), gpus,
                          _opts(), chunk=chunk)
        assert sink.order == list(range(N))
        assert all(np.array_equal(a, b) for a, b in zip(sink.frames, want)), (gpus, chunk)
        assert st.boundary_frames == st.chunks - 1


def test_process_per_gpu_matches_single_gpu(gpu_lib, tmp_path, monkeypatch):
    from video_restore_b200 import multiproc
    from video_restore_b200.cli import build_parser, config_from_args, frame_opts_from_config
    from video_restore_b200.pipeline import SyntheticSource

    g = _devices()
    monkeypatch.chdir(tmp_path)  # no models/ directory here: random-init weights, explicitly allowed
    args = build_parser().parse_args(["in", "out", "--model", NAME, "--quality", "fast", "--enhanced", "--tile-size", "64"])
    cfg = config_from_args(args)
    opts = frame_opts_from_config(cfg)
    total = 21
    src = SyntheticSource(48, 64, total, seed=1, distinct=8)
    from video_restore_b200.cli import make_restorer

    one = make_restorer(cfg, 0, allow_random=True)
    one.temporal_reset()
    want = {i: multiproc.frame_digest(one.process_frame(f, opts)) for i, f in enumerate(src.reader().read_range(0, total))}
    one.close()
    for world in sorted({2, g}):
        res = multiproc.run_job(list(range(world)), cfg, opts, total, True, height=48, width=64, allow_random=True)
        assert res["frames"] == total and res["per_frame"] == want, f"{world} processes"
        assert res["digest"] == multiproc.combine_digests(want)
        print(f"\n[multigpu] {world} processes: digest {res['digest']}, boundary exchange {res['exchange_ms']:.2f} ms")


def test_process_per_gpu_video_segments(gpu_lib, tmp_path, monkeypatch):
    """File in, one playable segment per GPU out; concatenated in rank order they equal the single-GPU output frame by frame
    (compared after the same lossy mp4v encode: decode both and compare)."""
    import cv2

    from video_restore_b200.cli import main

    _devices()
    monkeypatch.chdir(tmp_path)
    src = tmp_path / "in.mp4"
    wr = cv2.VideoWriter(str(src), cv2.VideoWriter_fourcc(*"mp4v"), 24.0, (64, 48))
    for i in range(10):
        wr.write(np.full((48, 64, 3), 20 + 25 * i, np.uint8))
    wr.release()
    common = ["--model", NAME, "--quality", "fast", "--enhanced", "--random-weights"]
    assert main([str(src), str(tmp_path / "one.mp4"), "--gpus", "0", *common]) == 0
    assert main([str(src), str(tmp_path / "two.mp4"), "--gpus", "0", "1", "--procs", *common]) == 0

    def frames_of(p):
        cap, out = cv2.VideoCapture(str(p)), []
        while True:
            ok, f = cap.read()
            if not ok:
                return out
            out.append(f)

    a = frames_of(tmp_path / "one.mp4")
    b = frames_of(tmp_path / "two.part00.mp4") + frames_of(tmp_path / "two.part01.mp4")
    assert len(a) == len(b) == 10
    # each segment is encoded on its own (its first frame is a key frame): same frames up to the codec's quantisation
    for x, y in zip(a, b):
        d = np.abs(x.astype(np.int16) - y.astype(np.int16))
        assert d.mean() < 1.0 and d.max() <= 12
