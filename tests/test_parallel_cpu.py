"""CPU, world_size 2 over gloo: host-side logic of the data-parallel path
(seed broadcast, in-place large-tensor all-reduce from grad hooks, coalesced
small tensors, averaging)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bayeformers_b200 as bf
    from bayeformers_b200 import parallel

    bf.manual_seed(1000 + rank)  # ranks start with different seeds
    seed = parallel.broadcast_seed(0)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(64, 64), torch.nn.Linear(64, 3))  # 4096-element weight = "large"
    sync = parallel.GradSync(net, large_numel=1024, bucketed=False)  # per-tensor mode
    x = torch.full((4, 64), float(rank + 1))
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    # hooks fired during backward: the large tensor is already in flight, small ones queued
    n_async, n_small = len(sync._handles), len(sync._small)
    sync.finish()
    out = dict(seed=seed, grads=[p.grad.clone() for p in net.parameters()], local=local, n_async=n_async,
               n_small=n_small, bytes=sync.bytes_last_step)
    lq, lp = parallel.all_reduce_elbo(torch.tensor([1.0 + rank]), torch.tensor([10.0 * (rank + 1)]))
    out["elbo"] = (float(lq), float(lp))
    # bucketed mode (the default when world > 1): gradients live in flat buffers, one all-reduce per bucket, launched
    # from the hook of the bucket's last tensor; two steps to exercise zero_grad / re-arming
    sync.remove()
    for p in net.parameters():
        p.grad = None
    frozen = torch.nn.Parameter(torch.ones(3))  # trainable-looking, never gets a gradient (a MOPED prior, quirk Q5)
    net.register_parameter("never_used", frozen)
    bs = parallel.GradSync(net, buckets=2, overlap=bool(rank >= 0 and os.environ.get("BF_TEST_OVERLAP") == "1"))
    assert bs.bucketed and not bs.buckets  # buckets are built by the first finish()
    steps, views = [], None
    for k in range(3):
        bs.zero_grad()
        net(x * (k + 1)).sum().backward()
        launched = sum(b.handle is not None for b in bs.buckets)
        bs.finish()
        if k == 0:
            views = [p.grad.data_ptr() for p in net.parameters() if p.grad is not None]
        steps.append(dict(grads=[p.grad.clone() for p in net.parameters() if p.grad is not None], launched=launched,
                          bytes=bs.bytes_last_step))
    out["bucketed"] = steps
    out["views_kept"] = views == [p.grad.data_ptr() for p in net.parameters() if p.grad is not None]
    out["n_buckets"] = len(bs.buckets)
    out["frozen_untouched"] = frozen.grad is None and id(frozen) not in bs._bucket_of
    ret[rank] = out
    dist.destroy_process_group()


def test_gradsync_world2_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
        assert r0["seed"] == r1["seed"] == 1000  # rank 0's seed everywhere -> identical Philox weights
        assert r0["n_async"] == 1 and r0["n_small"] == 3
        for g0, g1, l0, l1 in zip(r0["grads"], r1["grads"], r0["local"], r1["local"]):
            assert torch.equal(g0, g1)
            assert torch.allclose(g0, (l0 + l1) / 2)
        assert r0["bytes"] == sum(p.numel() * 4 for p in r0["grads"])
        assert r0["elbo"] == r1["elbo"] == (3.0, 30.0)
        # bucketed mode gives the same averaged gradients (step k uses x * (k+1): the gradient of the weights scales)
        assert r0["views_kept"] and r1["views_kept"] and r0["frozen_untouched"] and 1 <= r0["n_buckets"] <= 3
        for k in range(3):
            b0, b1 = r0["bucketed"][k], r1["bucketed"][k]
            assert b0["launched"] == 0  # default: nothing is launched during backward (see GradSync docstring)
            for g0, g1 in zip(b0["grads"], b1["grads"]):
                assert torch.equal(g0, g1)
            assert torch.allclose(b0["grads"][0], r0["grads"][0] * (k + 1))
            assert sum(p.numel() * 4 for p in r0["grads"]) <= b0["bytes"] <= sum(p.numel() * 4 for p in r0["grads"]) + 256


def test_gradsync_single_process_is_noop():
    from bayeformers_b200 import parallel

    net = torch.nn.Linear(4, 4)
    sync = parallel.GradSync(net)
    net(torch.ones(2, 4)).sum().backward()
    g = net.weight.grad.clone()
    sync.finish()
    assert torch.equal(g, net.weight.grad) and sync.world == 1


def _worker_samples(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bayeformers_b200 as bf
    from bayeformers_b200 import parallel, runtime

    bf.manual_seed(77 + rank)
    S = 4
    s_local = parallel.shard_samples(S)
    # stand-in for the per-sample logits of this rank's samples: a linear map with a shared parameter
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(3, 5))
    x = torch.arange(2 * s_local * 5, dtype=torch.float32).view(s_local, 2, 5) * (rank + 1) / 10
    raw = x @ w.t()                                   # [S_local, B=2, C=3]
    mean = parallel.mean_over_samples(raw, S)         # identical on both ranks
    loss = torch.nn.functional.cross_entropy(mean, torch.tensor([0, 2]))
    loss.backward()
    sync_grad = w.grad.clone()
    dist.all_reduce(sync_grad, op=dist.ReduceOp.SUM)  # what GradSync(average=False) does
    ret[rank] = dict(s_local=s_local, seed=runtime.seed(), mean=mean.detach().clone(), loss=float(loss),
                     grad=sync_grad, raw=raw.detach().clone(), x=x)
    dist.destroy_process_group()


def test_sample_sharding_world2_gloo():
    """shard_samples / mean_over_samples: two ranks holding 2 of 4 samples each reproduce the single-process
    loss and gradient of CE(mean over the 4 samples of the logits) after a SUM all-reduce of the gradients."""
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_samples, args=(world, port, ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
        assert r0["s_local"] == r1["s_local"] == 2
        assert r0["seed"] != r1["seed"]  # independent eps streams per rank
        assert torch.equal(r0["mean"], r1["mean"]) and r0["loss"] == r1["loss"]
        assert torch.equal(r0["grad"], r1["grad"])
        # single-process reference over all 4 samples
        torch.manual_seed(0)
        w = torch.nn.Parameter(torch.randn(3, 5))
        raw = torch.cat([r0["x"], r1["x"]]) @ w.t()
        loss = torch.nn.functional.cross_entropy(raw.mean(0), torch.tensor([0, 2]))
        loss.backward()
        assert abs(float(loss) - r0["loss"]) < 1e-6
        assert torch.allclose(w.grad, r0["grad"], rtol=1e-5, atol=1e-7)


def test_sample_sharding_single_process():
    from bayeformers_b200 import parallel

    assert parallel.shard_samples(8) == 8
    raw = torch.randn(4, 2, 3)
    assert torch.allclose(parallel.mean_over_samples(raw, 4), raw.mean(0))
