"""Frame-range sharding logic on the CPU: range arithmetic, and a world_size-2 gloo run whose output must be
bit-identical to the single-shard run (the temporal stage couples neighbouring frames across the boundary)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from video_restore_b200.sharder import FrameRangeSharder, shard_range  # noqa: E402

N_FRAMES, H, W = 7, 16, 20
MODEL = "RealESRGAN_x4_v3"


def test_shard_ranges_partition():
    for total in (0, 1, 7, 8, 3000, 3001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(3000, 3, 8) == (1125, 1500)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _make_oracle():
    from oracle.pipeline import OracleRestorer
    from util import oracle_model_from_sd, random_state_dict

    sd = random_state_dict(MODEL, seed=0)
    return OracleRestorer(MODEL, tile=12, tile_pad=2, model=oracle_model_from_sd(MODEL, sd))


def _frames():
    from video_restore_b200.synth import synth_frame

    return [synth_frame(H, W, seed=21, index=i) for i in range(N_FRAMES)]


def _opts():
    from oracle.pipeline import FrameOpts

    return FrameOpts(denoise=True, sharpen=0.3, clahe=True, temporal=True, temporal_tau=40.0)


def _worker(rank, world, port, out_dir, defer_head=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = _frames()
    results = {}
    sh = FrameRangeSharder(rank, world, N_FRAMES)
    orc = _make_oracle()
    calls = [0]
    inner = orc.upscale_only
    orc.upscale_only = lambda f, o: (calls.__setitem__(0, calls[0] + 1), inner(f, o))[1]
    from oracle import filters as OF
    if defer_head == "stream":
        # the pipelined host path of the product restorer (process_stream), stood in for by a generator over the oracle
        orc.process_stream = lambda fs, o: (orc.process_frame(f, o) for f in fs)
        n = sh.run_stream(orc, lambda i: frames[i], lambda i, o: results.__setitem__(i, o.copy()), _opts(),
                          temporal_blend=OF.temporal_blend)
    else:
        n = sh.run(orc, lambda i: frames[i], lambda i, o: results.__setitem__(i, o), _opts(), defer_head=defer_head,
                   temporal_blend=OF.temporal_blend)
    assert n == len(results)
    # the boundary protocol of step 1 costs one extra upscale per sending rank; the deferred head costs none
    assert calls[0] == n + (0 if defer_head or rank == world - 1 else 1)
    np.savez(Path(out_dir) / f"rank{rank}.npz", **{str(k): v for k, v in results.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,defer_head", [(2, False), (2, True), (3, True), (3, "stream")])
def test_two_shards_equal_one_shard(tmp_path, world, defer_head):
    torch.set_num_threads(2)
    frames = _frames()
    single = {}
    FrameRangeSharder(0, 1, N_FRAMES).run(_make_oracle(), lambda i: frames[i],
                                           lambda i, o: single.__setitem__(i, o), _opts())
    assert len(single) == N_FRAMES
    # the temporal stage really acts across the shard boundary (otherwise this test proves nothing)
    from oracle.pipeline import FrameOpts
    from dataclasses import replace
    orc = _make_oracle()
    no_t = orc.process_frame(frames[3], replace(_opts(), temporal=False))
    assert not np.array_equal(no_t, single[3])

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path), defer_head), nprocs=world, join=True)
    merged = {}
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        merged.update({int(k): z[k] for k in z.files})
    assert sorted(merged) == list(range(N_FRAMES))
    for i in range(N_FRAMES):
        assert np.array_equal(merged[i], single[i]), f"frame {i} differs between 2-shard and 1-shard runs"
