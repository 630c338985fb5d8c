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


def _ddp_worker(rank, world, port, out):
    """data-parallel training plumbing on CPU tensors: per-rank flat gradient arenas are summed bucket by bucket and the
    ranks' parameters stay identical after the (restated) Adam + EMA update with grad_scale = 1/world"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "potential-motion-plan-release_b200"))
    from pmp_b200 import train as ptrain
    from oracle import train_ref
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 10007
    g = torch.Generator().manual_seed(100 + rank)
    G = torch.randn(n, generator=g)
    local = G.clone()
    red = ptrain.GradAllReduce(G, bucket_mb=0.01)          # ~2,600 floats per bucket -> 4 buckets, back to front
    assert red.world == world and len(red.bounds) == 4 and red.bounds[0][1] == n and red.bounds[-1][0] == 0
    assert all(a[0] == b[1] for a, b in zip(red.bounds, red.bounds[1:]))
    red.run()
    P = torch.ones(n); M = torch.zeros(n); V = torch.zeros(n)
    train_ref.adam_step(P, G / world, M, V, 1, 1e-3)
    torch.save(dict(local=local, summed=G, P=P), out % rank)
    dist.destroy_process_group()


def test_two_rank_bucketed_gradient_allreduce(tmp_path):
    out = str(tmp_path / "ddp%d.pt")
    mp.spawn(_ddp_worker, args=(2, 29519, out), nprocs=2, join=True)
    r0, r1 = torch.load(out % 0), torch.load(out % 1)
    assert torch.equal(r0["summed"], r1["summed"])
    assert torch.allclose(r0["summed"], r0["local"] + r1["local"])
    assert torch.equal(r0["P"], r1["P"])


def test_flat_parameter_arena_is_laid_out_in_forward_module_order():
    """the reverse pass finishes gradients from the end of the arena: final conv last in the layout, embeddings first"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "potential-motion-plan-release_b200"))
    from pmp_b200 import train as ptrain
    from oracle import synth
    from diffuser.models import TemporalUnet_WCond
    model = TemporalUnet_WCond(**synth.arch("m2d")["unet"])
    before = {k: v.clone() for k, v in model.state_dict().items()}
    flat = ptrain.FlatParams(model, "cpu")
    ranks = [ptrain.FlatParams._rank(k) for k in flat.names]
    assert ranks == sorted(ranks) and ranks[0] == 0 and ranks[-1] == 5
    assert flat.names[-1].startswith("final_conv.") and flat.names[0].startswith("time_mlp.")
    assert all(o % 4 == 0 for o in flat.offsets) and flat.total >= sum(v.numel() for v in before.values() if v.dtype.is_floating_point) - 1
    # parameters are now views of the arena and unchanged in value; state_dict keeps the reference's key order
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k])
    assert list(model.state_dict().keys()) == list(before.keys())
    p = dict(model.named_parameters())[flat.names[3]]
    assert p.data_ptr() == flat.P[flat.offsets[3]:].data_ptr()
