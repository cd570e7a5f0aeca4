"""The N>1 path on CPU: world_size-2 gloo processes shard a batch of streams, each runs its shard (here on
the CPU oracle -- tests may use it as the stand-in compute), and the gathered result must equal the
single-process result.  Covers shard_bounds / gather_streams incl. uneven shards and gather-to-rank-0."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuralampmodelercore_b200 import sharding
from tests import nam_fixtures as fx


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 4096, 32768, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == sharding.max_shard(n, world) or n == 0
    with pytest.raises(ValueError):
        sharding.shard_bounds(8, 2, 2)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_streams: int, frames: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle

        x = fx.synthetic_batch(n_streams, frames, seed=77)
        lo, hi = sharding.shard_bounds(n_streams, world, rank)
        proto = oracle.OracleModel.from_dict(fx.load_model("wavenet"))
        proto.reset(48000.0, 64)
        y_local = torch.from_numpy(proto.run_batch(np.ascontiguousarray(x[lo:hi]), 64, 1))
        full = sharding.gather_streams(y_local, n_streams)
        on0 = sharding.gather_streams(y_local, n_streams, dst=0)
        dist.barrier()
        if rank == 0:
            q.put((full.numpy(), on0.numpy()))
        else:
            assert on0 is None
            q.put(("ok", full.shape))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_streams", [6, 5])
def test_two_rank_gloo_shard_and_gather(n_streams):
    frames, world = 300, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = next(r for r in results if not isinstance(r[0], str))
    from oracle import oracle

    proto = oracle.OracleModel.from_dict(fx.load_model("wavenet"))
    proto.reset(48000.0, 64)
    ref = proto.run_batch(fx.synthetic_batch(n_streams, frames, seed=77), 64, 1)
    assert np.array_equal(full[0], ref) and np.array_equal(full[1], ref)
