"""Contiguous frame-range sharding across GPUs with a single boundary-frame exchange (north star (4)).

Replaces the reference's round-robin multi-GPU dispatch (video_upscaler.py:430-488: `frame_idx % n_gpus` tag on a
shared queue, frames dequeued by the "wrong" worker are dropped, :471-473). Frames are independent except for the
temporal-consistency stage, which needs up_{t-1} -- the previous frame's UN-blended upscaled result. So:

    rank g owns frames [g*F/G, (g+1)*F/G)
    every rank walks its range with a RESET temporal state: its head frame comes out un-blended (u_first) and is held
    back; at the end of the range the rank posts a receive for rank g-1's last un-blended frame and sends its own to
    rank g+1 -- ALL boundaries in flight at once, one transfer of 3*sH*sW bytes each -- and finishes the head frame
    with one stand-alone temporal kernel. No collective, no barrier, no redundant upscale.

Because the temporal stage is non-recursive, the sharded result is bit-identical to the single-GPU run.

Transport. One process per GPU (torchrun); `torch.distributed` is plumbing only. With the nccl backend the boundary frame
never leaves device memory: `vr_temporal_get_prev(is_device=1)` -> one grouped isend/irecv pair over NVLink (receive posted
first, `batch_isend_irecv` = one ncclGroupStart/End, so the G-1 hops run concurrently instead of as a chain) ->
`vr_temporal_device` on the receiver. `connect()` opens the two point-to-point connections during set-up: NCCL creates them
lazily and that costs ~0.2 s per peer the first time (round 1 measured 228 / 584 / 1495 ms at 2 / 4 / 8 ranks for a chain of
blocking send-then-recv calls; the transfer itself is ~60 us). With gloo (CPU tests) the same protocol moves numpy arrays.
(In-process multi-GPU, one thread per GPU, is pipeline.py: there the boundary frame is one cudaMemcpyPeerAsync.)

`run(..., defer_head=False)` keeps the simpler protocol of SURVEY.md 8(e) for callers that want frames delivered strictly in
order: every rank except the last upscales its LAST frame first and sends it, at the price of one extra frame per shard.
"""
from __future__ import annotations

from typing import Callable

import numpy as np


def shard_range(total_frames: int, rank: int, world: int) -> tuple[int, int]:
    """[start, end) of `rank`'s contiguous range; ranges differ by at most one frame."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(total_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _backend() -> str:
    import torch.distributed as dist

    return dist.get_backend() if dist.is_available() and dist.is_initialized() else ""


class FrameRangeSharder:
    """Drives one restorer over its shard. `get_frame(i)` returns frame i (uint8 HxWx3 BGR); `put_frame(i, out)`
    receives the restored frame. `send(arr, dst)` / `recv(shape, src)` replace the transport (tests); by default it
    is torch.distributed point-to-point: device-resident with nccl, numpy with gloo."""

    def __init__(self, rank: int, world: int, total_frames: int, send: Callable | None = None,
                 recv: Callable | None = None):
        self.rank, self.world, self.total = rank, world, total_frames
        self.start, self.end = shard_range(total_frames, rank, world)
        self._send, self._recv = send, recv
        self._connected = False
        self.exchange_ms = 0.0   # host-side time of the last boundary exchange (transfer + head-frame blend)

    # -- transport ----------------------------------------------------------------------------------
    def _peers(self):
        return (self.rank - 1 if self.rank > 0 else None, self.rank + 1 if self.rank < self.world - 1 else None)

    def _p2p(self, send_t, recv_t) -> None:
        """One grouped exchange: receive from the left neighbour (posted first), send to the right one."""
        import torch.distributed as dist

        left, right = self._peers()
        ops = []
        if recv_t is not None and left is not None:
            ops.append(dist.P2POp(dist.irecv, recv_t, left))
        if send_t is not None and right is not None:
            ops.append(dist.P2POp(dist.isend, send_t, right))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def connect(self) -> None:
        """Open the point-to-point connections to both neighbours now (a few bytes each way), so that the boundary
        frame at the end of the range pays for the transfer only. No-op for a custom transport or a single shard."""
        if self._connected or self.world == 1 or self._send is not None or not _backend():
            return
        import torch

        dev = "cuda" if _backend() == "nccl" else "cpu"
        left, right = self._peers()
        self._p2p(torch.zeros(8, dtype=torch.uint8, device=dev) if right is not None else None,
                  torch.zeros(8, dtype=torch.uint8, device=dev) if left is not None else None)
        if dev == "cuda":
            torch.cuda.current_stream().synchronize()
        self._connected = True

    def _exchange_device(self, restorer, sH: int, sW: int):
        """nccl: this shard's last un-blended frame -> right neighbour, left neighbour's -> a device tensor (or None)."""
        import torch

        left, right = self._peers()
        send_t = recv_t = None
        if right is not None:
            send_t = torch.empty((sH, sW, 3), dtype=torch.uint8, device="cuda")
            restorer.temporal_get_prev(sH, sW, device_ptr=send_t.data_ptr())  # device-to-device, synchronises the stream
        if left is not None:
            recv_t = torch.empty((sH, sW, 3), dtype=torch.uint8, device="cuda")
        self._p2p(send_t, recv_t)
        torch.cuda.current_stream().synchronize()
        return recv_t

    def _exchange_host(self, arr, shape):
        """gloo / custom transport: numpy arrays."""
        left, right = self._peers()
        if self._send is not None:
            if right is not None:
                self._send(arr, right)
            return self._recv(shape, left) if left is not None else None
        import torch

        send_t = torch.from_numpy(np.ascontiguousarray(arr)) if right is not None else None
        recv_t = torch.empty(shape, dtype=torch.uint8) if left is not None else None
        self._p2p(send_t, recv_t)
        return recv_t.numpy() if recv_t is not None else None

    def _device_path(self, restorer) -> bool:
        return self._send is None and _backend() == "nccl" and hasattr(restorer, "temporal_blend_device")

    # -- protocol -----------------------------------------------------------------------------------
    def _finish_head(self, restorer, head: np.ndarray, opts, temporal_blend: Callable | None) -> np.ndarray:
        """End of the range: exchange the boundary frames (all shards at once) and blend this shard's head frame with the
        left neighbour's last un-blended frame."""
        import time

        t0 = time.perf_counter()
        sH, sW = head.shape[:2]
        if self._device_path(restorer):
            import torch

            prev = self._exchange_device(restorer, sH, sW)
            if prev is not None:
                d_head = torch.from_numpy(head).cuda()
                d_out = torch.empty_like(d_head)
                torch.cuda.current_stream().synchronize()
                restorer.temporal_blend_device(d_head.data_ptr(), prev.data_ptr(), sH, sW, d_out.data_ptr(),
                                               opts.temporal_alpha, opts.temporal_tau)
                restorer.sync()
                head = d_out.cpu().numpy()
        else:
            mine = restorer.temporal_get_prev(sH, sW) if self.rank < self.world - 1 else None
            prev = self._exchange_host(mine, head.shape)
            if prev is not None:
                if temporal_blend is None:
                    from .restorer import temporal_blend as _tb
                    temporal_blend = _tb
                head = temporal_blend(head, prev, opts.temporal_alpha, opts.temporal_tau)
        self.exchange_ms = (time.perf_counter() - t0) * 1e3
        return head

    def exchange_boundary(self, restorer, get_frame, opts) -> None:
        """The in-order protocol (defer_head=False): every rank but the last upscales its LAST frame first and hands the
        un-blended result to its right neighbour, which seeds its temporal state with it. No-op without the temporal stage
        or with a single shard."""
        if not opts.temporal or self.world == 1:
            return
        if self.end - self.start == 0:
            raise ValueError("empty shard: need total_frames >= world size when the temporal stage is on")
        self.connect()
        left, right = self._peers()
        f0 = get_frame(self.start)
        s = restorer.scale
        sH, sW = f0.shape[0] * s, f0.shape[1] * s
        if right is not None:
            restorer.temporal_reset()
            restorer.process_frame(get_frame(self.end - 1), opts)  # reset state: un-blended, and kept as the temporal state
        if self._device_path(restorer):
            prev = self._exchange_device(restorer, sH, sW)
            if prev is not None:
                restorer.temporal_set_prev(prev.data_ptr(), device_ptr=True, shape=(sH, sW))
        else:
            mine = restorer.temporal_get_prev(sH, sW) if right is not None else None
            prev = self._exchange_host(mine, (sH, sW, 3))
            if prev is not None:
                restorer.temporal_set_prev(prev)
        if left is None:
            restorer.temporal_reset()

    def run(self, restorer, get_frame, put_frame, opts, defer_head: bool = False, temporal_blend: Callable | None = None) -> int:
        """`temporal_blend(cur, prev, alpha, tau)` finishes a deferred head frame on the host transports; default: the
        CUDA stand-alone kernel (video_restore_b200.restorer.temporal_blend), tests pass the oracle's."""
        if not (defer_head and opts.temporal and self.world > 1):
            self.exchange_boundary(restorer, get_frame, opts)
            for i in range(self.start, self.end):
                put_frame(i, restorer.process_frame(get_frame(i), opts))
            return self.end - self.start
        n = self.end - self.start
        if n == 0:
            raise ValueError("empty shard: need total_frames >= world size when the temporal stage is on")
        self.connect()
        restorer.temporal_reset()
        head = restorer.process_frame(get_frame(self.start), opts)  # reset state: passes through un-blended
        for i in range(self.start + 1, self.end):
            put_frame(i, restorer.process_frame(get_frame(i), opts))
        put_frame(self.start, self._finish_head(restorer, head, opts, temporal_blend))
        return n

    def run_stream(self, restorer, get_frame, put_frame, opts, temporal_blend: Callable | None = None,
                   defer_head: bool = True) -> int:
        """As run() over the restorer's pipelined host path (process_stream: H2D / compute / D2H of neighbouring frames
        overlapped). Frames handed to put_frame are views of pinned ring buffers, valid until the next-but-two call.
        defer_head=False delivers strictly in order (a sequential encoder) at the price of one extra frame per shard."""
        n = self.end - self.start
        deferred = bool(opts.temporal) and self.world > 1 and defer_head
        if bool(opts.temporal) and self.world > 1 and not defer_head:
            self.exchange_boundary(restorer, get_frame, opts)
            for k, out in enumerate(restorer.process_stream((get_frame(i) for i in range(self.start, self.end)), opts)):
                put_frame(self.start + k, out)
            return n
        if deferred:
            if n == 0:
                raise ValueError("empty shard: need total_frames >= world size when the temporal stage is on")
            self.connect()
        restorer.temporal_reset()
        head = None
        for k, out in enumerate(restorer.process_stream((get_frame(i) for i in range(self.start, self.end)), opts)):
            if k == 0 and deferred:
                head = out.copy()  # the ring buffer behind `out` is reused three frames later
            else:
                put_frame(self.start + k, out)
        if deferred:
            put_frame(self.start, self._finish_head(restorer, head, opts, temporal_blend))
        return n
