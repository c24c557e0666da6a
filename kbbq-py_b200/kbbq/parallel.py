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
