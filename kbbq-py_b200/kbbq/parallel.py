"""Multi-GPU plumbing of the path: contiguous read shards, one integer all-reduce.

One process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).  Reads are
independent and the count tables are additive, so the only exchange step is a sum over the packed
int64 table buffer; every rank then recomputes the deltas from the same integers.
"""
import torch
import torch.distributed as dist


def shard_range(n_reads, rank, world, align=16):
    """Contiguous read range [lo, hi) of `rank`; boundaries are multiples of `align` reads so that
    every shard starts 16-byte aligned for any read length (super-rows never straddle ranks)."""
    units = (n_reads + align - 1) // align
    lo = min(n_reads, (units * rank // world) * align)
    hi = min(n_reads, (units * (rank + 1) // world) * align)
    return lo, hi


def allreduce_tables(tables, group=None):
    """In-place SUM of the packed [pos_errs | pos_total | din_errs | din_total] int64 buffer."""
    assert tables.dtype == torch.int64
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tables, op=dist.ReduceOp.SUM, group=group)
    return tables


def max_over_ranks(value, device, group=None):
    """Max of a python float over ranks (device-timed milliseconds in bench.py)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class _DevArray:
    """A device buffer owned by the C library, viewed by torch without a copy (__cuda_array_interface__)."""

    def __init__(self, ptr, n, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def session_tables_tensor(session, device):
    """The session's packed table buffer as an int64 torch tensor on `device` (no copy)."""
    ptr, n, _ = session.tables_dev()
    return torch.as_tensor(_DevArray(ptr, n), device=device)


def recalibrate_host_distributed(seq, qual, corr, rg, second, L, R, out, session=None, group=None, device=None,
                                 minscore=6):
    """One rank's share of the whole path under torch.distributed (one process per GPU):
    this rank's reads (host arrays, rg = global first-seen numbers) go through a session (pass 1), the partial
    tables of all ranks are summed in place with ONE all-reduce (NCCL over NVLink on GPUs) on the session's own
    table buffer, every rank recomputes the same deltas and applies them to its reads (pass 2) into `out`.
    Returns the session (reusable for the next call with the same shape)."""
    from . import _native
    n = seq.shape[0] if seq.ndim == 2 else seq.size // L
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if session is None:
        # this rank's share of the host threads (one node: every rank of the group runs on this host)
        import os
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        threads = max(1, len(os.sched_getaffinity(0)) // max(1, min(world, torch.cuda.device_count())))
        session = _native.Session(L, R, minscore, chunk_reads=0, resident_reads=max(n, 1), device=dev.index,
                                  host_threads=threads)
    else:
        session.reset()
    C = session.chunk_reads
    seq2, qual2, corr2 = seq.reshape(-1, L), qual.reshape(-1, L), corr.reshape(-1, L)
    # pass 1 over the whole shard in one call: the session pipelines the chunks (the copies of chunk k + 1 are
    # queued before the host cores pack chunk k)
    session.build_range(seq2[:n], qual2[:n], corr2[:n], None if rg is None else rg[:n], None if second is None else second[:n])
    session.flush()                                # the partial tables are complete
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        t = session_tables_tensor(session, dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        torch.cuda.current_stream(dev).synchronize()
    session.model()
    out2 = out.reshape(-1, L)
    for k, lo in enumerate(range(0, n, C)):
        session.apply_resident(k, out2[lo:min(n, lo + C)])
    session.sync()
    return session
