"""Ensemble sharding across the GPUs of one box and the end-of-run reductions.

Members are independent: rank r of G owns the contiguous block [r N / G, (r+1) N / G); the only
communication is the final gather of per-member fidelities / exit codes, or the all-reduce of a fidelity histogram.
"""
import numpy as np

from . import _lib


def shard_bounds(n_total, rank, world_size):
    """Contiguous block partition; the first n_total % world_size ranks get one extra member."""
    base, extra = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def fidelity_histogram(fidelity, lo=0.0, hi=1.0, nbins=256, counts=None):
    """Device histogram (m4q_hist_fidelity); counts are ADDED to `counts` (int64 CUDA tensor) if given."""
    t = _lib.require_cuda()
    f = fidelity.to(device='cuda', dtype=t.float64).contiguous()
    if counts is None:
        counts = t.zeros(nbins, dtype=t.int64, device='cuda')
    _lib.check(_lib.lib().m4q_hist_fidelity(f.numel(), _lib.ptr(f), float(lo), float(hi), int(nbins), _lib.ptr(counts),
                                            _lib.stream_ptr()))
    return counts


def gather_results(local, n_total, group=None):
    """All-gather variable-size per-rank arrays (torch tensors on the backend's device) into rank order."""
    import torch.distributed as dist
    t = _lib.torch()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank]
    width = max(sizes)
    pad = t.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [t.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return t.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def allreduce_histogram(counts, group=None):
    import torch.distributed as dist
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts
