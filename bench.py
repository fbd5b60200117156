#!/usr/bin/env python
"""bench.py -- frames/s of the per-frame restoration hot path on B200 (contract: see the task brief / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One "step" = one frame through the whole hot path (bilateral -> tiled Real-ESRGAN -> blend -> unsharp -> CLAHE ->
temporal). Default workload = BASELINE.json configs[3]: RealESRGAN_x4plus 720p -> 2880p, `--quality max --enhanced`
preset of the reference CLI (tile 512, overlap 64, video_upscaler.py:690-691) with every enhancement on.
  value : frames/s with frames resident in HBM (vr_restore_device_async), CUDA events on the library's stream
  e2e   : frames/s through FrameRestorer.process_stream with pinned HOST buffers (H2D + D2H of every frame inside the
          timed region, overlapped with the compute of neighbouring frames)
  roofline : conv kernel (K1) -- executed conv FLOPs of the step / summed conv-kernel time, vs MEASURED_PEAKS.json
  cpu_baseline : the oracle (CPU fp32 restatement of the reference path) on this box's host cores, bounded sample
`--impl reference` times only that CPU path (rank 0), printing the same metric.
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


def run_cpu_reference(wl, crop, steps, warmup):
    """The reference-equivalent PyTorch CPU path (oracle) on a crop x crop sample of the workload's frame.
    Returns (frames/s equivalent, seconds per sample step, description)."""
    import torch

    from oracle.pipeline import FrameOpts as OOpts
    from oracle.pipeline import OracleRestorer
    from video_restore_b200.models import MODEL_ZOO, flops_per_input_pixel
    from video_restore_b200.synth import random_state_dict, synth_frame

    sys.path.insert(0, str(ROOT / "tests"))
    from util import oracle_model_from_sd

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = random_state_dict(wl["model"], seed=0)
    model = oracle_model_from_sd(wl["model"], sd)
    orc = OracleRestorer(wl["model"], tile=wl["tile"], tile_pad=wl["pad"], blend=wl["blend"], model=model)
    opts = OOpts(**wl["opts"])
    ch, cw = min(crop, wl["H"]), min(crop, wl["W"])
    full = [synth_frame(wl["H"], wl["W"], seed=11, index=i) for i in range(2)]
    frames = [np.ascontiguousarray(f[:ch, :cw]) for f in full]
    for i in range(max(warmup, 1)):
        orc.process_frame(frames[i % 2], opts)
    t0 = time.perf_counter()
    for i in range(steps):
        orc.process_frame(frames[i % 2], opts)
    dt = (time.perf_counter() - t0) / steps
    per_px = flops_per_input_pixel(MODEL_ZOO[wl["model"]])
    sample_flops = per_px * ch * cw  # per_px is per frame pixel for every model (x2plus included)
    _, executed = conv_flops(wl)
    cpu_tflops = sample_flops / dt / 1e12
    fps = cpu_tflops * 1e12 / executed
    desc = (f"{ch}x{cw} crop of the {wl['H']}x{wl['W']} frame through the same chain (one tile), {steps} timed steps of "
            f"{dt:.2f} s; scaled to a full frame by executed conv FLOPs ({executed / 1e12:.2f} TFLOP/frame, "
            f"CPU ran {cpu_tflops:.3f} TFLOP/s)")
    return fps, dt, desc, cores


def main():
    _saved_stdout_fd = None
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-crop", type=int, default=None, help="edge of the CPU-baseline sample crop")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        crop = args.cpu_crop or 160
        fps, dt, desc, cores = run_cpu_reference(wl, crop, max(args.steps, 1), max(args.warmup, 1))
        line = {"impl": "reference", "metric": metric, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch

    from video_restore_b200.restorer import FrameOpts, FrameRestorer
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

    K, Wm = args.steps, max(args.warmup, 3)
    H, W = wl["H"], wl["W"]
    s = 2 if "x2" in wl["model"] else 4
    opts = FrameOpts(**wl["opts"])
    sd = random_state_dict(wl["model"], seed=0)
    r = FrameRestorer(wl["model"], sd, tile=wl["tile"], tile_pad=wl["pad"], blend=wl["blend"], gpu_id=local_rank)
    # this rank's contiguous frame range: [rank*K, (rank+1)*K); a few distinct frames, generated on the host
    n_src = 4
    host_frames = [synth_frame(H, W, seed=11, index=rank * K + i) for i in range(n_src)]
    pinned_in = [torch.from_numpy(f).pin_memory() for f in host_frames]
    d_in = [t.cuda(non_blocking=False) for t in pinned_in]
    d_out = torch.empty((H * s, W * s, 3), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.ExternalStream(r.stream)

    # ---- boundary-frame exchange of the shard protocol (once per shard; timed on its own) ----
    boundary_ms = 0.0
    if world > 1 and opts.temporal:
        from dataclasses import replace
        barrier()
        t0 = time.perf_counter()
        if rank < world - 1:
            r.process_frame_device(d_in[(K - 1) % n_src].data_ptr(), H, W, d_out.data_ptr(), replace(opts, temporal=False))
            torch.cuda.synchronize()
            dist.send(d_out, dst=rank + 1)
        if rank > 0:
            prev = torch.empty_like(d_out)
            dist.recv(prev, src=rank - 1)
            torch.cuda.synchronize()
            r.temporal_set_prev(prev.data_ptr(), device_ptr=True, shape=(H * s, W * s))
        torch.cuda.synchronize()
        boundary_ms = (time.perf_counter() - t0) * 1e3

    # ---- device-resident throughput ----
    for i in range(Wm):
        r.process_frame_device(d_in[i % n_src].data_ptr(), H, W, d_out.data_ptr(), opts, sync=True)
    clocks = ClockSampler(local_rank)
    launches0 = r.launch_count
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    conv_ms_total = 0.0
    e0.record(stream)
    for i in range(K):
        r.process_frame_device(d_in[i % n_src].data_ptr(), H, W, d_out.data_ptr(), opts, sync=False)
    e1.record(stream)
    r.sync()
    barrier()
    clk = clocks.stop()
    launches = r.launch_count - launches0
    dev_ms = e0.elapsed_time(e1)
    # conv-kernel time of one representative step (events inside the library bracket each tile's network)
    r.process_frame_device(d_in[0].data_ptr(), H, W, d_out.data_ptr(), opts, sync=True)
    total_ms_1, conv_ms_1 = r.last_timing()

    # ---- end to end through the public API: host frames in, host frames out ----
    # FrameRestorer.process_stream (vr_submit / vr_wait): every frame is copied host -> device from pinned memory,
    # restored, and copied device -> host into pinned memory, all inside the timed region; copies of neighbouring
    # frames overlap the compute (two frames in flight). The consumer reads one pixel of every output frame.
    in_np = [t.numpy() for t in pinned_in]
    for _ in r.process_stream((in_np[i % n_src] for i in range(3)), opts):
        pass
    barrier()
    t0 = time.perf_counter()
    checksum = 0
    for out in r.process_stream((in_np[i % n_src] for i in range(K)), opts):
        checksum += int(out[0, 0, 0])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    if dist is not None:
        t = torch.tensor([dev_ms, e2e_s * 1e3, boundary_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, boundary_ms = [float(x) for x in t.tolist()]
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    else:
        e2e_ms = e2e_s * 1e3
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = load_peaks()
    k1_traffic = {}
    # per-launch DRAM bytes of the dominant conv kernel from the committed ncu capture (K3; the K1-era record is kept)
    for tp in (ROOT / "profiles" / "k3_traffic.json", ROOT / "profiles" / "k1_traffic.json"):
        if tp.exists():
            k1_traffic = json.loads(tp.read_text())
            break
    from video_restore_b200.models import MODEL_ZOO, conv_layers
    n_conv_launches = len(conv_layers(MODEL_ZOO[wl["model"]])) + (1 if MODEL_ZOO[wl["model"]]["kind"] == "rrdb" else 0)
    fps = world * K / (dev_ms / 1e3)
    e2e_fps = world * K / (e2e_ms / 1e3)
    conv_tflops = executed / (conv_ms_1 / 1e3) / 1e12
    hbm_side = None
    if k1_traffic.get("dram_bytes_per_launch") and wl["model"] == "RealESRGAN_x4plus":
        # the ncu capture is the 720p single-tile x4plus frame: bytes scale with the pixels the launches process
        from video_restore_b200.models import flops_per_input_pixel
        scale = executed / (flops_per_input_pixel(MODEL_ZOO[wl["model"]]) * 720 * 1280)
        gbs = k1_traffic["dram_bytes_per_launch"] * scale * n_conv_launches / (conv_ms_1 / 1e3) / 1e9
        hbm_side = {"achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                    "note": "DRAM bytes per launch from the ncu capture (720p single tile) x pixel ratio x launches / conv time"}
    line = {
        "metric": metric, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic", "config": config,
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": H * W * 3,
                "d2h_bytes_per_step": H * s * W * s * 3, "ms_per_step": e2e_ms / K},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "conv3x3_pair_kernel (K3: 348 of 360 conv launches at x4plus; K1 runs the rest)", "achieved": conv_tflops,
                     "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": conv_tflops / peaks["sustained"],
                     "peak_burst": peaks["burst"], "frac_of_burst": conv_tflops / peaks["burst"],
                     "peak_source": peaks["source"] + " (sustained figure: kernel timed inside a long step)",
                     "traffic": k1_traffic.get("dram_bytes_per_launch"), "traffic_source": k1_traffic.get("source"),
                     "launches_per_step": n_conv_launches,
                     "algorithmic_flop_per_launch": executed / max(n_conv_launches, 1),
                     "avg_launch_us": conv_ms_1 * 1e3 / max(n_conv_launches, 1),
                     "conv_ms_per_step": conv_ms_1, "step_ms": total_ms_1,
                     # second roofline of the same launches: DRAM bytes (ncu, per launch) x launches / conv time against the measured
                     # copy bandwidth -- the 32-channel layers of K3 sit on this one
                     "hbm": hbm_side,
                     "useful_tflops_whole_step": useful / (dev_ms / K / 1e3) / 1e12,
                     "useful_frac_of_sustained": useful / (dev_ms / K / 1e3) / 1e12 / peaks["sustained"]},
        "boundary_exchange_ms": boundary_ms,
    }
    if world == 1 and not args.no_cpu_baseline:
        crop = args.cpu_crop or 256
        cfps, cdt, desc, cores = run_cpu_reference(wl, crop, 2, 1)
        line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc}
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
