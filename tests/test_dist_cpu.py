"""N > 1 host logic on CPU: two gloo ranks shard the reads, sum their partial tables with the
product's all-reduce helper and must reproduce the single-process tables bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, N, L, R, q):
    sys.path.insert(0, PKG)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle  # the checker stands in for the kernels: this test is about the exchange step
    from kbbq import parallel, synth
    lo, hi = parallel.shard_range(N, rank, world)
    seq, qual, corr, rg, second = synth.synth_reads(99, lo, hi - lo, L, R)  # each rank regenerates its own shard
    pe, pt, de, dt = oracle.build_tables(seq, qual, corr, rg, second, L, R)
    packed = torch.from_numpy(np.concatenate([a.ravel() for a in (pe, pt, de, dt)]))
    parallel.allreduce_tables(packed)
    t_ms = parallel.max_over_ranks(1.0 + rank, "cpu")
    if rank == 0:
        q.put((packed.numpy().copy(), t_ms))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_tables_equal_single_process(world, oracle_mod):
    from kbbq import synth
    N, L, R = 4001, 50, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, L, R, q)) for r in range(world)]
    for p in procs:
        p.start()
    packed, t_ms = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seq, qual, corr, rg, second = synth.synth_reads(99, 0, N, L, R)
    want = np.concatenate([a.ravel() for a in oracle_mod.build_tables(seq, qual, corr, rg, second, L, R)])
    assert np.array_equal(packed, want)
    assert t_ms == float(world)  # max over ranks
