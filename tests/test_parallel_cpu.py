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
    sync = parallel.GradSync(net, large_numel=1024)
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


def test_gradsync_single_process_is_noop():
    from bayeformers_b200 import parallel

    net = torch.nn.Linear(4, 4)
    sync = parallel.GradSync(net)
    net(torch.ones(2, 4)).sum().backward()
    g = net.weight.grad.clone()
    sync.finish()
    assert torch.equal(g, net.weight.grad) and sync.world == 1
