"""CPU, world_size 2 over gloo: the N>1 plumbing (contiguous problem shards, final gather,
max-over-ranks timing) without a GPU."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_prob, n_samp, out):
    sys.path.insert(0, os.path.join(ROOT, "potential-motion-plan-release_b200"))
    from pmp_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_problems(n_prob, n_samp, rank, world)
    # stand-in for the per-rank planning result: row i carries its global index
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 3)
    full = shard.gather_rows(local.reshape(-1, n_samp, 3).reshape(-1, 3), n_prob * n_samp, world) \
        if False else None
    # gather per problem (rows of n_samp candidates stay together)
    per_prob = local.reshape(-1, n_samp * 3)
    full = shard.gather_rows(per_prob, n_prob, world).reshape(-1, 3)
    slowest = shard.max_over_ranks(1.0 + rank, "cpu")
    if rank == 0:
        torch.save(dict(full=full, slowest=slowest), out)
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    sys.path.insert(0, os.path.join(ROOT, "potential-motion-plan-release_b200"))
    from pmp_b200 import shard
    for n in (0, 1, 7, 8, 10000):
        for w in (1, 2, 3, 8):
            r = [shard.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    assert shard.shard_problems(5, 10, 1, 2) == (30, 50)


def test_two_rank_gather(tmp_path):
    out = str(tmp_path / "r0.pt")
    n_prob, n_samp = 7, 4
    mp.spawn(_worker, args=(2, 29517, n_prob, n_samp, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["slowest"] == 2.0
    assert torch.equal(res["full"][:, 0], torch.arange(n_prob * n_samp, dtype=torch.float32))
