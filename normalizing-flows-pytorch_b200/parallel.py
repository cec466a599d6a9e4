"""Sample-sharded evaluation across the GPUs of one box.

Every layer is per-sample independent in eval mode, so rank r simply owns rows [r*B/N, (r+1)*B/N) of the global
batch with replicated weights; there is NO data-path collective.  The only exchange is one all-reduce (SUM) of the
2-element fp64 payload (sum of per-sample NLL, sample count) -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist

from .likelihood import LN2


def shard_bounds(n_rows, rank, world_size):
    """Contiguous, balanced row range of `rank`: the first n_rows % world_size ranks get one extra row."""
    base, extra = divmod(n_rows, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(x, rank, world_size):
    lo, hi = shard_bounds(x.size(0), rank, world_size)
    return x[lo:hi]


def allreduce_nll(total, group=None):
    """In-place SUM all-reduce of the (sum NLL, count) payload; no-op without an initialised process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return total


def global_bits_per_dim(total, D, group=None):
    total = allreduce_nll(total.clone(), group)
    s, n = (float(v) for v in total.tolist())
    return s / n / (D * LN2)


def sharded_bits_per_dim(model, x_local, group=None):
    """bits/dim of the GLOBAL batch given this rank's shard `x_local` (already on this rank's GPU)."""
    _, total = model.nll(x_local)
    return global_bits_per_dim(total, x_local[0].numel(), group)
