#!/usr/bin/env python
"""bench.py -- frames/s of the per-frame restoration hot path on B200 (contract: see the task brief / DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference] [--frames F]

One "step" = one frame through the whole hot path (bilateral -> tiled Real-ESRGAN -> blend -> unsharp -> CLAHE ->
temporal). Default workload = BASELINE.json configs[3]: RealESRGAN_x4plus 720p -> 2880p, `--quality max --enhanced`
preset of the reference CLI (tile 512, overlap 64, video_upscaler.py:690-691) with every enhancement on.
  value     : frames/s with frames resident in HBM (vr_restore_device_async), CUDA events on the library's stream. With N > 1
              every rank runs a contiguous frame range and the boundary-frame exchange of the shard protocol (one grouped
              NCCL isend/irecv per rank over NVLink + one temporal kernel on the head frame) is INSIDE the timed region.
  e2e       : frames/s through the public host API (FrameRangeSharder.run_stream -> FrameRestorer.process_stream) with pinned
              HOST buffers: H2D + D2H of every frame inside the timed region, overlapped with neighbouring frames' compute
  roofline  : the conv kernels -- conv FLOPs of a step / conv-kernel time AVERAGED OVER THE TIMED STEPS (per-frame CUDA
              events inside the library), vs MEASURED_PEAKS.json; `frac_executed` counts the padded tiles the reference's
              tile loop makes every implementation compute, `frac_useful` only the unpadded frame; `frac` =
              frac_executed (the kernel's own efficiency)
  tolerance : the step's result vs the CPU oracle on the FULL frame (max LSB, PSNR), outside the timed region
  cpu_baseline : the oracle (CPU fp32 restatement of the reference path) on this box's host cores: ONE real full frame
`--impl reference` times only that CPU path (rank 0): every step is one padded tile of the frame's real tile grid through
the fp32 network (the whole frame's tiles are cycled), the rest of the chain is timed once; nothing is scaled by FLOPs.
`--frames F` = strong scaling: F frames in total split into contiguous ranges over the N ranks (BASELINE configs[4]).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ALL_ON = dict(denoise=True, sharpen=0.5, clahe=True, temporal=True)
WORKLOADS = {
    # name: model, LR frame, tile/pad, blend, enhancement opts
    "c4_x4plus_720p_qmax_enhanced": dict(model="RealESRGAN_x4plus", H=720, W=1280, tile=512, pad=64,
                                         blend="gaussian", opts=ALL_ON),
    "c4_x4plus_720p_qmax_plain": dict(model="RealESRGAN_x4plus", H=720, W=1280, tile=1536, pad=10, blend="crop",
                                      opts={}),
    "c1_x4plus_256_tile128": dict(model="RealESRGAN_x4plus", H=256, W=256, tile=128, pad=16, blend="crop", opts={}),
    "c2_x4v3_480p_fast": dict(model="RealESRGAN_x4_v3", H=480, W=854, tile=1024, pad=10, blend="crop", opts={}),
    "c3_x2plus_1080p_seamless": dict(model="RealESRGAN_x2plus", H=1080, W=1920, tile=512, pad=32, blend="gaussian",
                                     opts={}),
    "c5_x4plus_1080p": dict(model="RealESRGAN_x4plus", H=1080, W=1920, tile=1024, pad=10, blend="crop", opts={}),
    # configs[4] with the temporal stage on, so that the frame-range shards really exchange their boundary frame
    "c5_x4plus_1080p_temporal": dict(model="RealESRGAN_x4plus", H=1080, W=1920, tile=1024, pad=10, blend="crop",
                                     opts=dict(temporal=True)),
}
DEFAULT_WORKLOAD = "c4_x4plus_720p_qmax_enhanced"
METRIC = {"RealESRGAN_x4plus": "x4plus 720p->2880p frames/sec"}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(burst=float(d["bf16_tflops"]), sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), source="MEASURED_PEAKS.json")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def conv_flops(wl):
    """(useful, executed) conv FLOPs per frame. executed = sum over the reference's padded tiles."""
    from video_restore_b200.models import MODEL_ZOO, flops_per_input_pixel

    spec = MODEL_ZOO[wl["model"]]
    per_px = flops_per_input_pixel(spec)
    H, W, tile, pad = wl["H"], wl["W"], wl["tile"], wl["pad"]
    if spec["scale"] == 2:
        H, W = H + H % 2, W + W % 2
    useful = per_px * H * W
    executed = 0
    for y in range(0, H, tile):
        for x in range(0, W, tile):
            x1, y1 = min(x + tile, W), min(y + tile, H)
            pw = min(x1 + pad, W) - max(x - pad, 0)
            ph = min(y1 + pad, H) - max(y - pad, 0)
            executed += per_px * pw * ph
    return float(useful), float(executed)


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksEventReasonSwPowerCap if hasattr(nv, "nvmlClocksEventReasonSwPowerCap")
                 else nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU oracle legs (the only places this file touches oracle/): cpu_baseline + tolerance, and --impl reference
# ----------------------------------------------------------------------------------------------------------------------
def _oracle_restorer(wl, model=None):
    import torch

    from oracle.pipeline import OracleRestorer
    from video_restore_b200.synth import random_state_dict

    sys.path.insert(0, str(ROOT / "tests"))
    from util import oracle_model_from_sd

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if model is None:
        model = oracle_model_from_sd(wl["model"], random_state_dict(wl["model"], seed=0))
    return OracleRestorer(wl["model"], tile=wl["tile"], tile_pad=wl["pad"], blend=wl["blend"], model=model), cores


def cpu_full_frame(wl, frame, prev_up=None):
    """ONE real full frame through the whole oracle chain, stage by stage. Returns dict(seconds, up = the frame after the
    upscale stage (bilateral -> tiled network -> merge/blend), out = end of the chain, cores)."""
    from oracle import filters as OF
    from oracle.pipeline import FrameOpts as OOpts

    orc, cores = _oracle_restorer(wl)
    o = OOpts(**wl["opts"])
    orc.process_frame(np.ascontiguousarray(frame[:32, :32]))  # thread pool / allocator warm-up on a 32x32 corner
    t0 = time.perf_counter()
    f = OF.bilateral_filter(frame, o.denoise_d, o.denoise_sigma_color, o.denoise_sigma_space) if o.denoise else frame
    up, _ = orc.upsampler.enhance(f, outscale=orc.scale)
    t_up = time.perf_counter()
    out = up
    if o.sharpen > 0:
        out = OF.unsharp_mask(out, o.sharpen)
    if o.clahe:
        out = OF.clahe_bgr(out, o.clahe_clip, o.clahe_grid)
    if o.temporal:
        out = OF.temporal_blend(out, prev_up if prev_up is not None else out, o.temporal_alpha, o.temporal_tau)
    t1 = time.perf_counter()
    return dict(seconds=t1 - t0, seconds_upscale=t_up - t0, up=up, out=out, cores=cores)


def cpu_crop_sample(wl, crop, steps, warmup):
    """Fallback sample for `--cpu-crop`: a crop x crop corner of the frame as ONE tile, scaled to the frame by executed conv
    FLOPs -- an EXTRAPOLATION (kind "port-extrapolated"); the default legs time real tiles / a real full frame instead."""
    from oracle.pipeline import FrameOpts as OOpts
    from video_restore_b200.models import MODEL_ZOO, flops_per_input_pixel
    from video_restore_b200.synth import synth_frame

    orc, cores = _oracle_restorer(wl)
    opts = OOpts(**wl["opts"])
    ch, cw = min(crop, wl["H"]), min(crop, wl["W"])
    frames = [np.ascontiguousarray(synth_frame(wl["H"], wl["W"], seed=11, index=i)[:ch, :cw]) for i in range(2)]
    for i in range(max(warmup, 1)):
        orc.process_frame(frames[i % 2], opts)
    t0 = time.perf_counter()
    for i in range(steps):
        orc.process_frame(frames[i % 2], opts)
    dt = (time.perf_counter() - t0) / steps
    per_px = flops_per_input_pixel(MODEL_ZOO[wl["model"]])
    _, executed = conv_flops(wl)
    cpu_tflops = per_px * ch * cw / dt / 1e12
    fps = cpu_tflops * 1e12 / executed
    desc = (f"EXTRAPOLATED: {ch}x{cw} crop of the {wl['H']}x{wl['W']} frame as one tile, {steps} timed steps of {dt:.2f} s, "
            f"scaled to a full frame by executed conv FLOPs ({executed / 1e12:.2f} TFLOP/frame; CPU ran {cpu_tflops:.3f} TFLOP/s)")
    return fps, dt, desc, cores


def cpu_tile_steps(wl, steps, warmup):
    """The reference arm's sample: every step = one padded tile of the frame's REAL tile grid through the fp32 network (tiles
    cycled in grid order, so with steps >= tiles every tile of the frame is timed at least once); the rest of the chain
    (bilateral, merge / Gaussian blend, unsharp, CLAHE, temporal, conversions) is timed once on the full frame with the
    network stubbed out. frame time = sum over tiles of that tile's mean step time + rest. Nothing is scaled by FLOPs."""
    import torch

    from oracle.pipeline import FrameOpts as OOpts
    from oracle.realesrganer import tile_grid
    from video_restore_b200.synth import synth_frame

    orc, cores = _oracle_restorer(wl)
    s = orc.scale
    H, W = wl["H"], wl["W"]
    Hp, Wp = (H + H % 2, W + W % 2) if s == 2 else (H, W)
    grid = tile_grid(Hp, Wp, wl["tile"], wl["pad"], s).tolist()
    frame = synth_frame(H, W, seed=11, index=0)
    x = torch.from_numpy(np.ascontiguousarray(frame[:, :, ::-1].astype(np.float32) / 255.0)).permute(2, 0, 1)[None]
    if (Hp, Wp) != (H, W):
        x = torch.nn.functional.pad(x, (0, Wp - W, 0, Hp - H), "reflect")
    tiles = [x[:, :, t[6]:t[7], t[4]:t[5]].contiguous() for t in grid]
    model = orc.model
    k = 0
    with torch.no_grad():
        for _ in range(max(warmup, 1)):
            model(tiles[k % len(tiles)])
            k += 1
        per_tile = [[] for _ in tiles]
        t_all0 = time.perf_counter()
        for _ in range(steps):
            i = k % len(tiles)
            t0 = time.perf_counter()
            model(tiles[i])
            per_tile[i].append(time.perf_counter() - t0)
            k += 1
        t_all = time.perf_counter() - t_all0
    timed = [np.mean(v) for v in per_tile if v]
    # tiles never reached (steps < tiles): priced at the mean seconds per tile pixel of the timed ones -- stated in `sample`
    px = [(t[7] - t[6]) * (t[5] - t[4]) for t in grid]
    sec_per_px = sum(np.mean(v) for v in per_tile if v) / max(sum(p for p, v in zip(px, per_tile) if v), 1)
    net_s = sum(np.mean(v) if v else sec_per_px * p for p, v in zip(px, per_tile))
    # the rest of the chain with the network replaced by a stub of the right output shape
    stub, _ = _oracle_restorer(wl, model=_StubNet(s))
    o = OOpts(**wl["opts"])
    if o.temporal:
        stub.temporal_set_prev(np.zeros((H * s, W * s, 3), np.uint8))
    t0 = time.perf_counter()
    stub.process_frame(frame, o)
    rest_s = time.perf_counter() - t0
    frame_s = net_s + rest_s
    missing = sum(1 for v in per_tile if not v)
    desc = (f"{steps} timed steps, each ONE padded tile of the frame's real {len(tiles)}-tile grid through the fp32 network "
            f"(tiles cycled; mean step {t_all / max(steps, 1):.2f} s); frame = sum of the per-tile means {net_s:.2f} s + the rest of "
            f"the chain timed once on the full frame {rest_s:.2f} s"
            + (f"; {missing} tile(s) not reached, priced per pixel" if missing else "; every tile timed"))
    return 1.0 / frame_s, t_all / max(steps, 1), desc, cores


class _StubNet:
    """Stands in for the network when timing everything BUT the network: zeros of the network's output shape."""

    def __init__(self, scale):
        self.scale = scale

    def eval(self):
        return self

    def __call__(self, x):
        import torch

        return torch.full((x.shape[0], 3, x.shape[2] * self.scale, x.shape[3] * self.scale), 0.5)


def tolerance_verdict(up_gpu, chain_gpu, chain_from_gpu_up, cpu):
    """max LSB / PSNR of the upscale stage over the FULL frame vs the oracle; the enhancement stage is checked bit for bit on the
    GPU's own upscaled frame (CLAHE amplifies +-1 LSB input differences, so its end-to-end figure is reported, not gated)."""
    sys.path.insert(0, str(ROOT / "tests"))
    from util import max_lsb, psnr_u8, psnr_unsaturated

    ref = cpu["up"]
    lsb, p = max_lsb(up_gpu, ref), psnr_u8(up_gpu, ref)
    pu, frac = psnr_unsaturated(up_gpu, ref)
    v = {"max_lsb": lsb, "psnr_db": round(p, 2), "psnr_db_unclamped": round(pu, 2), "unclamped_frac": round(frac, 4),
         "differing_frac": round(float((up_gpu != ref).mean()), 5), "frame": "full",
         "stage": "bilateral -> tiled network -> merge/blend (uint8 frame before the enhancement filters)",
         "bar": "max 1 LSB and PSNR >= 50 dB vs the fp32 CPU oracle"}
    ok = lsb <= 1 and p >= 50.0 and pu >= 50.0
    if chain_gpu is not None:
        v["enhancement_chain_bit_exact"] = bool(np.array_equal(chain_gpu, chain_from_gpu_up))
        v["chain_psnr_db_vs_oracle_chain"] = round(psnr_u8(chain_gpu, cpu["out"]), 2)
        ok = ok and v["enhancement_chain_bit_exact"]
    v["ok"] = bool(ok)
    return v


def main():
    _saved_stdout_fd = None
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=0, help="strong scaling: total frames split over the ranks")
    ap.add_argument("--cpu-crop", type=int, default=None, help="CPU legs on a crop (extrapolated) instead of real tiles / a full frame")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (no cpu_baseline, no tolerance)")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    metric = METRIC.get(wl["model"], f"{wl['model']} frames/sec")
    useful, executed = conv_flops(wl)
    config = {"workload": args.workload, "model": wl["model"], "frame": f"{wl['W']}x{wl['H']}",
              "tile": wl["tile"], "tile_pad": wl["pad"], "blend": wl["blend"], "enhance": wl["opts"],
              "weights": "random-init seed 0", "parallelism": f"frame-range shards x{world}",
              "l2": "per-layer working set (>=127 MB NHWC activations x3 buffers) exceeds the 126 MB L2; no flush",
              "conv_tflop_per_frame_useful": round(useful / 1e12, 3),
              "conv_tflop_per_frame_executed": round(executed / 1e12, 3)}

    # ------------------------------------------------------------------ reference arm (CPU oracle)
    if args.impl == "reference":
        if rank != 0:
            return 0
        if args.cpu_crop:
            fps, dt, desc, cores = cpu_crop_sample(wl, args.cpu_crop, max(args.steps, 1), max(args.warmup, 1))
            kind = "port-extrapolated"
        else:
            fps, dt, desc, cores = cpu_tile_steps(wl, max(args.steps, 1), max(args.warmup, 1))
            kind = "port"
        line = {"impl": "reference", "metric": metric, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": desc},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch

    from video_restore_b200.restorer import FrameOpts, FrameRestorer
    from video_restore_b200.sharder import FrameRangeSharder, shard_range
    from video_restore_b200.synth import random_state_dict, synth_frame

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the restoration path is CUDA-only (no CPU fallback)"}))
        return 2
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # native libraries (NCCL's version banner) write to fd 1: keep stdout for the one JSON line
        sys.stdout.flush()
        _saved_stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    Wm = max(args.warmup, 3)
    strong = args.frames > 0
    if strong:
        f_lo, f_hi = shard_range(args.frames, rank, world)
        K = f_hi - f_lo
        first_index = f_lo
    else:
        K = args.steps
        first_index = rank * K       # weak scaling: this rank's contiguous frame range [rank*K, (rank+1)*K)
    H, W = wl["H"], wl["W"]
    s = 2 if "x2" in wl["model"] else 4
    sH, sW = H * s, W * s
    opts = FrameOpts(**wl["opts"])
    sd = random_state_dict(wl["model"], seed=0)
    r = FrameRestorer(wl["model"], sd, tile=wl["tile"], tile_pad=wl["pad"], blend=wl["blend"], gpu_id=local_rank)
    # a few distinct frames of the range, generated on the host (3000 HR frames would be 298 GB: never materialised)
    n_src = 4
    host_frames = [synth_frame(H, W, seed=11, index=first_index + i) for i in range(n_src)]
    pinned_in = [torch.from_numpy(f).pin_memory() for f in host_frames]
    d_in = [t.cuda(non_blocking=False) for t in pinned_in]
    d_out = torch.empty((sH, sW, 3), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.ExternalStream(r.stream)
    deferred = world > 1 and bool(opts.temporal)   # the shard protocol's boundary exchange (sharder.py)
    d_head = torch.empty_like(d_out) if deferred else None
    d_send = torch.empty_like(d_out) if deferred and rank < world - 1 else None
    d_recv = torch.empty_like(d_out) if deferred and rank > 0 else None
    exch = {"ms": 0.0}

    def run_shard(n_frames):
        """This rank's frame range on device-resident frames; with N > 1 and the temporal stage on, ends with the boundary
        exchange: last un-blended frame -> rank+1 (device to device over NVLink, all ranks at once), head frame blended."""
        r.temporal_reset()
        for i in range(n_frames):
            dst = d_head if (deferred and i == 0) else d_out
            r.process_frame_device(d_in[i % n_src].data_ptr(), H, W, dst.data_ptr(), opts, sync=False)
        if deferred and n_frames > 0:
            if d_send is not None:
                r.temporal_get_prev(sH, sW, device_ptr=d_send.data_ptr())   # device-to-device copy, then stream sync
            else:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            ops = []
            if d_recv is not None:
                ops.append(dist.P2POp(dist.irecv, d_recv, rank - 1))       # receive posted first
            if d_send is not None:
                ops.append(dist.P2POp(dist.isend, d_send, rank + 1))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            torch.cuda.current_stream().synchronize()
            if d_recv is not None:
                r.temporal_blend_device(d_head.data_ptr(), d_recv.data_ptr(), sH, sW, d_out.data_ptr(),
                                        opts.temporal_alpha, opts.temporal_tau)
                torch.cuda.synchronize()
            exch["ms"] = (time.perf_counter() - t0) * 1e3

    # ---- device-resident throughput ----
    for i in range(Wm):
        r.process_frame_device(d_in[i % n_src].data_ptr(), H, W, d_out.data_ptr(), opts, sync=True)
    run_shard(min(Wm, 2))          # opens the NCCL point-to-point connections (lazy: ~0.2 s per peer the first time)
    r.sync()
    clocks = ClockSampler(local_rank)
    launches0, conv_launches0 = r.launch_count, r.conv_launch_count
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run_shard(K)
    e1.record(stream)
    r.sync()                        # also averages the per-frame conv / total events over the K frames just run
    barrier()
    clk = clocks.stop()
    launches = r.launch_count - launches0
    conv_launches_per_step = (r.conv_launch_count - conv_launches0) / max(K, 1)
    dev_ms = e0.elapsed_time(e1)
    total_ms_avg, conv_ms_avg = r.last_timing()
    timed_frames = r.last_timing_frames()
    boundary_ms = exch["ms"]   # in situ: includes waiting for a slower neighbour to reach the end of ITS range (rank skew)
    # the transfer on its own: every rank enters together (barrier), same buffers, same grouped isend / irecv
    transfer_ms = 0.0
    if deferred:
        barrier()
        t0 = time.perf_counter()
        ops = []
        if d_recv is not None:
            ops.append(dist.P2POp(dist.irecv, d_recv, rank - 1))
        if d_send is not None:
            ops.append(dist.P2POp(dist.isend, d_send, rank + 1))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        torch.cuda.current_stream().synchronize()
        transfer_ms = (time.perf_counter() - t0) * 1e3

    # ---- end to end through the public API: host frames in, host frames out ----
    # FrameRangeSharder.run_stream -> FrameRestorer.process_stream (vr_submit / vr_wait): every frame is copied host -> device
    # from pinned memory, restored, and copied device -> host into pinned memory, all inside the timed region; copies of
    # neighbouring frames overlap the compute (two frames in flight); with N > 1 the boundary exchange is inside as well.
    e2e_ms = None
    if not args.no_e2e:
        in_np = [t.numpy() for t in pinned_in]
        sh = FrameRangeSharder(rank, world, world * 3)
        sh.run_stream(r, lambda i: in_np[i % n_src], lambda i, o: None, opts)
        n_total = args.frames if strong else world * K
        sh = FrameRangeSharder(rank, world, n_total)
        sh._connected = True
        checksum = [0]
        barrier()
        t0 = time.perf_counter()
        sh.run_stream(r, lambda i: in_np[i % n_src], lambda i, o: checksum.__setitem__(0, checksum[0] + int(o[0, 0, 0])), opts)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        barrier()

    if dist is not None:
        t = torch.tensor([dev_ms, e2e_ms or 0.0, boundary_ms, conv_ms_avg, total_ms_avg, transfer_ms], device="cuda",
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_max, boundary_ms, conv_ms_avg, total_ms_avg, transfer_ms = [float(x) for x in t.tolist()]
        e2e_ms = e2e_max if e2e_ms is not None else None
        lt = torch.tensor([launches, K], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches, frames_all = int(lt[0].item()), int(lt[1].item())
    else:
        frames_all = K
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = load_peaks()
    k_traffic = {}
    # per-launch DRAM bytes of the dominant conv kernel from the committed ncu capture (newest record first)
    for tp in (ROOT / "profiles" / "k4_traffic.json", ROOT / "profiles" / "k3_traffic.json", ROOT / "profiles" / "k1_traffic.json"):
        if tp.exists():
            k_traffic = json.loads(tp.read_text())
            break
    from video_restore_b200.models import MODEL_ZOO, flops_per_input_pixel
    n_conv_launches = conv_launches_per_step   # counted by the library (rank 0's shard)
    fps = frames_all / (dev_ms / 1e3)
    conv_tflops = executed / (conv_ms_avg / 1e3) / 1e12
    conv_tflops_useful = useful / (conv_ms_avg / 1e3) / 1e12
    hbm_side = None
    per_frame = k_traffic.get("dram_bytes_per_frame_720p") or (
        k_traffic.get("dram_bytes_per_launch", 0) * k_traffic.get("launches", 0) / max(k_traffic.get("share_of_frame", 1.0), 1e-9))
    if per_frame and wl["model"] == "RealESRGAN_x4plus":
        # the ncu capture is the 720p single-tile x4plus frame: bytes scale with the pixels the launches process
        scale = executed / (flops_per_input_pixel(MODEL_ZOO[wl["model"]]) * 720 * 1280)
        gbs = per_frame * scale / (conv_ms_avg / 1e3) / 1e9
        hbm_side = {"achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                    "note": "DRAM bytes of all conv launches of a frame from the ncu capture (720p single tile) x pixel ratio / conv time"}
    line = {
        "metric": metric, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": dev_ms / max(K, 1), "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": config,
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"bound": "tensor",
                     "kernel": k_traffic.get("kernel", "conv3x3_pair_kernel (K3) and the other conv kernels of the frame"),
                     "achieved": conv_tflops, "peak": peaks["sustained"], "unit": "TFLOP/s",
                     "frac": conv_tflops / peaks["sustained"],
                     "frac_executed": conv_tflops / peaks["sustained"],
                     "frac_useful": conv_tflops_useful / peaks["sustained"],
                     "achieved_useful": conv_tflops_useful,
                     "peak_burst": peaks["burst"], "frac_executed_of_burst": conv_tflops / peaks["burst"],
                     "frac_useful_of_burst": conv_tflops_useful / peaks["burst"],
                     "peak_source": peaks["source"] + " (sustained figure: kernels timed inside a long step)",
                     "traffic": k_traffic.get("dram_bytes_per_launch"), "traffic_source": k_traffic.get("source"),
                     "launches_per_step": n_conv_launches,
                     "algorithmic_flop_per_launch": executed / max(n_conv_launches, 1),
                     "avg_launch_us": conv_ms_avg * 1e3 / max(n_conv_launches, 1),
                     "conv_ms_per_step": conv_ms_avg, "step_ms": total_ms_avg,
                     "timed_over": f"per-frame CUDA events averaged over {timed_frames} timed steps (max over ranks)",
                     "executed_vs_useful": "executed = padded tiles of the reference's tile loop (parity requires them); "
                                           "useful = unpadded frame (SURVEY 8(d))",
                     "hbm": hbm_side,
                     "useful_tflops_whole_step": useful / (dev_ms / max(K, 1) / 1e3) / 1e12},
        "boundary_exchange_ms": boundary_ms,
        "boundary_transfer_ms": transfer_ms,
        "multi_gpu": {"boundary_exchange_ms": boundary_ms, "boundary_transfer_ms": transfer_ms,
                      "note": "exchange = in situ at the end of each rank's range, inside the timed region: transfer + head-frame "
                              "blend + waiting for a slower left neighbour (rank skew); transfer = the same grouped isend/irecv "
                              "entered by all ranks together after a barrier (max over ranks)",
                      "inside_timed_region": bool(deferred),
                      "protocol": "contiguous frame ranges; last un-blended frame -> rank+1 by one grouped NCCL isend/irecv "
                                  "(device to device, all ranks concurrently) + one temporal kernel on the head frame"
                                  if deferred else "no exchange (single shard or temporal stage off)",
                      "frames_total": frames_all},
    }
    config["boundary_exchange_ms"] = round(boundary_ms, 3)
    config["boundary_transfer_ms"] = round(transfer_ms, 3)
    if e2e_ms is not None:
        n_e2e = args.frames if strong else world * K
        line["e2e"] = {"value": n_e2e / (e2e_ms / 1e3), "unit": "frames/s", "h2d_bytes_per_step": H * W * 3,
                       "d2h_bytes_per_step": sH * sW * 3, "ms_per_step": e2e_ms / max(K, 1)}
    if world == 1 and not args.no_cpu_baseline:
        if args.cpu_crop:
            cfps, cdt, desc, cores = cpu_crop_sample(wl, args.cpu_crop, 2, 1)
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port-extrapolated", "sample": desc}
            line["tolerance"] = None
        else:
            # GPU results of one frame (outside the timed region): upscale stage alone, and the whole chain on two frames
            from dataclasses import replace
            f0, f1 = host_frames[0], host_frames[1]
            up_opts = replace(FrameOpts(), denoise=opts.denoise, denoise_d=opts.denoise_d,
                              denoise_sigma_color=opts.denoise_sigma_color, denoise_sigma_space=opts.denoise_sigma_space)
            up1_gpu = r.process_frame(f1, up_opts)
            chain_gpu = chain_ref = prev_e = None
            post = opts.sharpen > 0 or opts.clahe or opts.temporal
            cpu_prev = None
            if post:
                from oracle import filters as OF

                def enh(u):
                    u = OF.unsharp_mask(u, opts.sharpen) if opts.sharpen > 0 else u
                    return OF.clahe_bgr(u, opts.clahe_clip, opts.clahe_grid) if opts.clahe else u

                up0_gpu = r.process_frame(f0, up_opts)
                r.temporal_reset()
                r.process_frame(f0, opts)
                chain_gpu = r.process_frame(f1, opts)
                prev_e = enh(up0_gpu)
                chain_ref = enh(up1_gpu)
                if opts.temporal:
                    chain_ref = OF.temporal_blend(chain_ref, prev_e, opts.temporal_alpha, opts.temporal_tau)
                cpu_prev = prev_e
            cpu = cpu_full_frame(wl, f1, prev_up=cpu_prev)
            line["cpu_baseline"] = {"value": 1.0 / cpu["seconds"], "unit": "frames/s", "cores": cpu["cores"], "kind": "port",
                                    "sample": f"ONE real full {W}x{H} frame through the whole oracle chain (all {os.cpu_count()} host "
                                              f"threads): {cpu['seconds']:.1f} s, of which the upscale stage {cpu['seconds_upscale']:.1f} s"}
            line["tolerance"] = tolerance_verdict(up1_gpu, chain_gpu, chain_ref, cpu)
    if _saved_stdout_fd is not None:
        sys.stdout.flush()
        os.dup2(_saved_stdout_fd, 1)
        os.close(_saved_stdout_fd)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
