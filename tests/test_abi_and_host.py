"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/vrb200.h
declares, the ctypes table matches, error behaviour without a GPU, and the CLI preset table."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "vrb200.h").read_text()


def declared_symbols():
    return sorted(set(re.findall(r"\b(vr_[a-z0-9_]+)\s*\(", HEADER)))


def test_library_exports_every_declared_symbol():
    from video_restore_b200 import _lib

    lib = _lib.load()  # raises if the .so is missing: there is no fallback
    names = declared_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vrb200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table and header disagree"


def test_header_cites_reference_interfaces():
    for cite in ("video_upscaler.py:328-338", "video_upscaler.py:490-505", "video_upscaler.py:496",
                 "video_upscaler.py:501", "video_upscaler.py:326"):
        assert cite in HEADER


def test_struct_layouts_match_header():
    from video_restore_b200 import _lib

    assert ctypes.sizeof(_lib.VrConfig) == 16 * 4
    assert ctypes.sizeof(_lib.VrFrameOpts) == 16 * 4


def test_no_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this is the CPU-box behaviour")
    from video_restore_b200._lib import VrError
    from video_restore_b200.restorer import FrameRestorer, bilateral_filter
    import numpy as np

    with pytest.raises(VrError, match="no CUDA device|no CPU fallback"):
        FrameRestorer("RealESRGAN_x4_v3", None)
    with pytest.raises(VrError):
        bilateral_filter(np.zeros((8, 8, 3), np.uint8))


def test_invalid_configs_rejected_before_touching_a_device():
    from video_restore_b200 import _lib

    lib = _lib.load()
    h = ctypes.c_void_p()
    for bad in (dict(model_kind=7), dict(scale=3), dict(tile=0), dict(pre_pad=10), dict(num_feat=48)):
        kw = dict(model_kind=0, scale=4, num_block=23, num_conv=0, num_feat=64, num_grow_ch=32, tile=512, tile_pad=10,
                  pre_pad=0, blend=0, device=0)
        kw.update(bad)
        cfg = _lib.VrConfig(**kw)
        assert lib.vr_create(ctypes.byref(cfg), ctypes.byref(h)) == -1, bad
        assert (lib.vr_last_error(None) or b"") != b""
    assert lib.vr_tile_grid(0, 10, 8, 2, 4, None, 0) == -1


def test_product_never_imports_oracle():
    for py in (ROOT / "video_restore_b200").glob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, py
    for cu in (ROOT / "video_restore_b200" / "csrc").glob("*"):
        if cu.is_file():
            assert "oracle/" not in cu.read_text(errors="ignore").replace("oracle/filters.py", "").replace(
                "oracle.realesrganer", "").replace("spec oracle", "").replace("/ oracle", ""), cu


# --- CLI presets: the table at reference video_upscaler.py:687-701 -------------------------------------------
@pytest.mark.parametrize("argv,tile,overlap,pad,crf,preset", [
    ([], 1024, 16, 10, 15, "slow"),
    (["--enhanced"], 512, 32, 32, 15, "slow"),
    (["--quality", "max"], 1536, 32, 10, 12, "veryslow"),
    (["--quality", "max", "--enhanced"], 512, 64, 64, 12, "veryslow"),
    (["--quality", "fast"], 1024, 16, 10, 18, "fast"),
    (["--quality", "fast", "--enhanced", "--tile-size", "256", "--tile-overlap", "8"], 256, 8, 8, 18, "fast"),
])
def test_cli_presets(argv, tile, overlap, pad, crf, preset):
    from video_restore_b200.cli import build_parser, config_from_args

    cfg = config_from_args(build_parser().parse_args(["in.mp4", "out.mp4", *argv]))
    assert (cfg.tile_size, cfg.tile_overlap, cfg.tile_pad, cfg.crf, cfg.preset) == (tile, overlap, pad, crf, preset)
    assert cfg.light_denoise == cfg.enhanced_mode and cfg.use_fp16


def test_cli_enhancement_flags():
    from video_restore_b200.cli import build_parser, config_from_args, frame_opts_from_config

    p = build_parser()
    cfg = config_from_args(p.parse_args(["a", "b", "--enhanced"]))
    o = frame_opts_from_config(cfg)
    assert o.denoise and (o.denoise_d, o.denoise_sigma_color, o.denoise_sigma_space) == (5, 25.0, 25.0)
    assert o.clahe and o.temporal and cfg.seamless and o.sharpen == pytest.approx(0.1)
    cfg = config_from_args(p.parse_args(["a", "b", "--enhanced", "--no-seamless", "--no-temporal",
                                         "--no-color-enhance", "--sharpen", "0.3", "--denoise", "0.3"]))
    o = frame_opts_from_config(cfg)
    assert not (cfg.seamless or o.temporal or o.clahe) and o.sharpen == pytest.approx(0.3)
    assert o.denoise_sigma_color == pytest.approx(50.0)
    cfg = config_from_args(p.parse_args(["a", "b"]))
    o = frame_opts_from_config(cfg)
    assert not (o.denoise or o.clahe or o.temporal or cfg.seamless) and o.sharpen == 0.0
    assert config_from_args(p.parse_args(["a", "b", "--model", "RealESRGAN_x2plus"])).scale == 2
    with pytest.raises(SystemExit):
        p.parse_args(["a", "b", "--model", "nope"])


def test_cli_missing_checkpoint_is_an_error_unless_random_weights_are_requested(tmp_path, monkeypatch, capsys):
    """the CLI used to fall back to random-init weights silently and write a full-length video of garbage. The
    reference downloads the checkpoint or fails (video_upscaler.py:342-367)."""
    from video_restore_b200.cli import MissingWeights, OptimizedConfig, build_parser, load_weights

    monkeypatch.chdir(tmp_path)   # no models/ directory here
    cfg = OptimizedConfig(model_name="RealESRGAN_x4_v3")
    with pytest.raises(MissingWeights, match="--random-weights"):
        load_weights(cfg)
    sd = load_weights(cfg, allow_random=True)
    assert "body.0.weight" in sd and "not a restored video" in capsys.readouterr().out
    args = build_parser().parse_args(["a", "b", "--random-weights", "--procs"])
    assert args.random_weights and args.procs


def test_frame_digests_are_order_defined():
    import numpy as np

    from video_restore_b200.multiproc import combine_digests, frame_digest

    a = np.arange(64 * 32 * 3, dtype=np.uint8).reshape(64, 32, 3)
    b = a.copy()
    b[8, 3, 1] ^= 1            # row 8 is one of the sampled rows (every 8th)
    assert frame_digest(a) != frame_digest(b)
    d1 = {0: frame_digest(a), 1: frame_digest(b)}
    d2 = {1: frame_digest(b), 0: frame_digest(a)}      # arrival order of the ranks does not matter
    assert combine_digests(d1) == combine_digests(d2) != combine_digests({0: d1[1], 1: d1[0]})


def _fake_ffmpeg(tmp_path, body: str):
    """A stand-in `ffmpeg` executable: records its argv, then runs `body` (python) with `argv` in scope."""
    import stat
    import sys

    exe = tmp_path / "ffmpeg"
    exe.write_text(f"#!{sys.executable}\nimport sys, json\nargv = sys.argv[1:]\n"
                   f"open({str(tmp_path / 'argv.json')!r}, 'w').write(json.dumps(argv))\n{body}\n")
    exe.chmod(exe.stat().st_mode | stat.S_IEXEC)
    return str(exe)


def test_copy_audio_muxes_through_ffmpeg_and_keeps_the_video_on_failure(tmp_path):
    """cli.copy_audio = the reference's _copy_audio (video_upscaler.py:604-627): video of the output + audio of the input,
    both copied, through a temp file that replaces the output; any failure leaves the output as written."""
    import json

    from video_restore_b200.cli import copy_audio

    src, dst = tmp_path / "in.mp4", tmp_path / "out.mp4"
    src.write_bytes(b"source-with-audio")
    dst.write_bytes(b"video-only")
    ok_bin = _fake_ffmpeg(tmp_path, "open(argv[-1], 'wb').write(b'muxed')")
    assert copy_audio(str(src), str(dst), ffmpeg_bin=ok_bin) is True
    assert dst.read_bytes() == b"muxed"
    argv = json.loads((tmp_path / "argv.json").read_text())
    i = [k for k, a in enumerate(argv) if a == "-i"]
    assert [argv[k + 1] for k in i] == [str(dst), str(src)]          # input 0 = our video, input 1 = the source
    assert argv[argv.index("-map") + 1] == "0:v" and "1:a" in argv   # video from ours, audio from the source
    assert argv[argv.index("-c:v") + 1] == "copy" and argv[argv.index("-c:a") + 1] == "copy"
    assert argv[-1] == str(dst) + ".temp.mp4" and not (tmp_path / "out.mp4.temp.mp4").exists()
    # no audio track / ffmpeg error: non-zero exit -> output untouched, temp removed
    dst.write_bytes(b"video-only")
    bad_bin = _fake_ffmpeg(tmp_path, "open(argv[-1], 'wb').write(b'partial'); sys.exit(1)")
    assert copy_audio(str(src), str(dst), ffmpeg_bin=bad_bin) is False
    assert dst.read_bytes() == b"video-only" and not (tmp_path / "out.mp4.temp.mp4").exists()
    # no binary at all: nothing happens
    assert copy_audio(str(src), str(dst), ffmpeg_bin=str(tmp_path / "absent")) is False
    assert dst.read_bytes() == b"video-only"


def test_join_segments_uses_the_concat_demuxer_and_removes_the_parts(tmp_path):
    import json

    from video_restore_b200.cli import join_segments

    parts = [tmp_path / f"out.part{i}.mp4" for i in range(3)]
    for i, p in enumerate(parts):
        p.write_bytes(b"seg%d" % i)
    out = tmp_path / "out.mp4"
    body = ("lst = argv[argv.index('-i') + 1]\n"
            "names = [l.split(\"'\")[1] for l in open(lst).read().splitlines()]\n"
            "open(argv[-1], 'wb').write(b''.join(open(n, 'rb').read() for n in names))")
    assert join_segments([str(p) for p in parts], str(out), ffmpeg_bin=_fake_ffmpeg(tmp_path, body)) is True
    assert out.read_bytes() == b"seg0seg1seg2" and not any(p.exists() for p in parts)
    argv = json.loads((tmp_path / "argv.json").read_text())
    assert argv[argv.index("-f") + 1] == "concat" and argv[argv.index("-c") + 1] == "copy"
    # failure keeps the parts
    for i, p in enumerate(parts):
        p.write_bytes(b"seg%d" % i)
    assert join_segments([str(p) for p in parts], str(out), ffmpeg_bin=_fake_ffmpeg(tmp_path, "sys.exit(2)")) is False
    assert all(p.exists() for p in parts)
