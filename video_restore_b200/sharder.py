"""Contiguous frame-range sharding across GPUs with a single boundary-frame exchange (north star (4)).

Replaces the reference's round-robin multi-GPU dispatch (video_upscaler.py:430-488: `frame_idx % n_gpus` tag on a
shared queue, frames dequeued by the "wrong" worker are dropped, :471-473). Frames are independent except for the
temporal-consistency stage, which needs up_{t-1} -- the previous frame's UN-blended upscaled result. So:

    rank g owns frames [g*F/G, (g+1)*F/G)
    1. every rank except the last upscales its LAST frame first (temporal off) and sends that 3*sH*sW-byte uint8
       frame to rank g+1 (one point-to-point transfer; device-to-device over NVLink with the nccl backend);
    2. every rank except the first receives it and seeds its temporal state with it;
    3. each rank then walks its range in order. No collective and no barrier on the per-frame path.

Because the temporal stage is non-recursive, the sharded result is bit-identical to the single-GPU run.
One process per GPU (torchrun); `torch.distributed` is plumbing only (send/recv of one frame per shard).

`run(..., defer_head=True)` is the variant without the redundant upscale of step 1 (the one pipeline.py uses in-process):
every rank walks its range with a RESET temporal state, so its head frame comes out un-blended (u_first) and is held back;
at the end it sends its last un-blended frame (the restorer's temporal state) to rank g+1, receives rank g-1's, and finishes
the head frame with one stand-alone temporal blend. Same single transfer per boundary, same bits, one frame less work per
shard; the head frame of a shard is delivered last (put_frame is called out of order for it).
"""
from __future__ import annotations

from dataclasses import replace
from typing import Callable

import numpy as np


def shard_range(total_frames: int, rank: int, world: int) -> tuple[int, int]:
    """[start, end) of `rank`'s contiguous range; ranges differ by at most one frame."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(total_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class FrameRangeSharder:
    """Drives one restorer over its shard. `get_frame(i)` returns frame i (uint8 HxWx3 BGR);
    `put_frame(i, out)` receives the restored frame. `send`/`recv` move one uint8 array between neighbouring
    ranks; by default they are torch.distributed point-to-point calls (gloo on CPU in tests, nccl on GPUs)."""

    def __init__(self, rank: int, world: int, total_frames: int, send: Callable | None = None,
                 recv: Callable | None = None):
        self.rank, self.world, self.total = rank, world, total_frames
        self.start, self.end = shard_range(total_frames, rank, world)
        self._send = send or self._dist_send
        self._recv = recv or self._dist_recv

    # -- default transport ------------------------------------------------------------------------
    @staticmethod
    def _dist_send(arr: np.ndarray, dst: int) -> None:
        import torch
        import torch.distributed as dist

        t = torch.from_numpy(np.ascontiguousarray(arr))
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.send(t, dst=dst)

    @staticmethod
    def _dist_recv(shape, src: int) -> np.ndarray:
        import torch
        import torch.distributed as dist

        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.empty(shape, dtype=torch.uint8, device=dev)
        dist.recv(t, src=src)
        return t.cpu().numpy()

    # -- protocol ---------------------------------------------------------------------------------
    def exchange_boundary(self, restorer, get_frame, opts) -> None:
        """Steps 1 and 2. No-op without the temporal stage or with a single shard. Empty shards forward nothing
        (ranges are non-empty whenever total_frames >= world)."""
        if not opts.temporal or self.world == 1:
            return
        if self.end - self.start == 0:
            raise ValueError("empty shard: need total_frames >= world size when the temporal stage is on")
        if self.rank < self.world - 1:
            up_last = restorer.process_frame(get_frame(self.end - 1), replace(opts, temporal=False))
            self._send(up_last, self.rank + 1)
        if self.rank > 0:
            f0 = get_frame(self.start)
            s = restorer.scale
            prev = self._recv((f0.shape[0] * s, f0.shape[1] * s, 3), self.rank - 1)
            restorer.temporal_set_prev(prev)
        else:
            restorer.temporal_reset()

    def run(self, restorer, get_frame, put_frame, opts, defer_head: bool = False, temporal_blend: Callable | None = None) -> int:
        """`temporal_blend(cur, prev, alpha, tau)` finishes a deferred head frame; default: the CUDA stand-alone kernel
        (video_restore_b200.restorer.temporal_blend), tests pass the oracle's."""
        if not (defer_head and opts.temporal and self.world > 1):
            self.exchange_boundary(restorer, get_frame, opts)
            for i in range(self.start, self.end):
                put_frame(i, restorer.process_frame(get_frame(i), opts))
            return self.end - self.start
        n = self.end - self.start
        if n == 0:
            raise ValueError("empty shard: need total_frames >= world size when the temporal stage is on")
        restorer.temporal_reset()
        head = restorer.process_frame(get_frame(self.start), opts)  # reset state: passes through un-blended
        for i in range(self.start + 1, self.end):
            put_frame(i, restorer.process_frame(get_frame(i), opts))
        if self.rank < self.world - 1:
            self._send(restorer.temporal_get_prev(head.shape[0], head.shape[1]), self.rank + 1)
        if self.rank > 0:
            prev = self._recv(head.shape, self.rank - 1)
            if temporal_blend is None:
                from .restorer import temporal_blend as _tb
                temporal_blend = _tb
            head = temporal_blend(head, prev, opts.temporal_alpha, opts.temporal_tau)
        put_frame(self.start, head)
        return n
