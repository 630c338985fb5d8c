"""Multi-GPU plumbing: independent planning problems are sharded across ranks in contiguous
blocks (all `samples_perprob` candidates of a problem stay on one rank so first-valid
selection is local); no collective on the data path, one gather of the results at the end
(SURVEY.md section 8e).  Works with any torch.distributed backend (nccl on B200s, gloo in
the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """contiguous [lo, hi) of rank; the first (n % world) ranks take one extra item"""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_problems(n_prob, n_samp, rank, world):
    """row range [lo, hi) of the [n_prob*n_samp] candidate batch owned by `rank`"""
    lo, hi = shard_range(n_prob, rank, world)
    return lo * n_samp, hi * n_samp


def gather_rows(local, n_total_rows, world=None, group=None):
    """all_gather of row blocks of unequal length: local [n_local, ...] -> [n_total_rows, ...] on every rank"""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = world or dist.get_world_size(group)
    counts = [shard_range(n_total_rows, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in counts)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, counts)], 0)


def max_over_ranks(value, device):
    """max of a python float over ranks (timings are reported as the slowest rank)"""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
