"""Two-rank NCCL worker of tests/test_gpu_models.py::test_two_rank_nccl_batch_and_sample_sharding
(also runnable by hand: `python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1
tests/multirank_worker.py`).  One process per GPU.  Checks, on real GPUs over NCCL:

  1. batch sharding: after `broadcast_seed` both ranks draw BIT-IDENTICAL weights with no communication;
  2. batch sharding: the all-reduced gradients (`GradSync`, average) equal the mean of the gradients the two batch
     shards give in one process, and log q / log p are identical on both ranks without any reduction;
  3. sample sharding: `shard_samples` + `mean_over_samples` + `GradSync(average=False)` + `all_reduce_elbo` reproduce the
     single-process S-sample step (pattern of /root/reference/examples/bert_squad.py:190-212 under DataParallel, :245).
Every rank asserts; rank 0 prints "MULTIRANK OK".
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("HF_HUB_OFFLINE", "1")

import torch
import torch.distributed as dist

import bayeformers_b200 as bf
from bayeformers_b200 import parallel, runtime

N_BATCHES = 100


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def build(dev):
    from transformers import BertConfig, BertForSequenceClassification
    torch.manual_seed(0)
    cfg = BertConfig(vocab_size=128, hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                     intermediate_size=1024, max_position_embeddings=64, num_labels=3)
    model = BertForSequenceClassification(cfg)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("bias"):
                p.add_(torch.randn(p.shape, generator=g) * 0.02)
    bm = bf.to_bayesian(model, delta=0.05, freeze=True, gemm_dtype="bf16", kl_grad=True)
    bf.accelerate_host_(bm, layernorm=True, fuse_gelu=True, fuse_residual=True)
    bm = bm.to(dev).eval()  # eval: the host model's torch dropout would draw rank-dependent masks
    bf.enable_presample(bm)
    bf.cast_frequentist_(bm, torch.bfloat16)
    return bm, cfg


def local_step(bm, ids, labels, S, S_total=None, group_mean=False):
    """forward + loss + backward of one rank (or of the single-process emulation); returns (loss, lq, lp)."""
    B = ids.shape[0]
    with bf.mc_samples(S):
        logits = bm(input_ids=ids.repeat(S, 1)).logits
    raw = logits.float().view(S, B, -1)
    lq, lp = bm.log_variational_posterior(), bm.log_prior()
    if group_mean:
        mean = parallel.mean_over_samples(raw, S_total)
        kl = (lq.sum() - lp.sum()) / S_total
    else:
        mean = raw.mean(0)
        kl = lq.mean() - lp.mean()
    loss = kl / N_BATCHES + torch.nn.functional.cross_entropy(mean, labels)
    return loss, lq.detach(), lp.detach()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    assert world == 2
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    bf.manual_seed(5000 + rank)  # ranks start apart
    bm, cfg = build(dev)
    params = [p for p in bm.parameters() if p.requires_grad]
    seed = parallel.broadcast_seed(0)
    assert seed == 5000
    S, B, Tn = 4, 8, 32
    g = torch.Generator().manual_seed(7)
    ids_all = torch.randint(0, cfg.vocab_size, (B, Tn), generator=g).to(dev)
    labels_all = torch.randint(0, 3, (B,), generator=g).to(dev)

    # ---- 1. identical weights on both ranks, no communication
    snap = bf.rng_state(bm)
    assert bm._presampler.run(S)
    sums = []
    for layer in bm._presampler.layers:
        W = layer._presampled[1]
        sums.append(W.view(torch.int16 if W.dtype == torch.bfloat16 else torch.int32).to(torch.int64).sum())
        layer._presampled = None
    sums = torch.stack(sums)
    gathered = [torch.empty_like(sums) for _ in range(world)]
    dist.all_gather(gathered, sums)
    assert torch.equal(gathered[0], gathered[1]), "ranks drew different weights"

    # ---- 2. batch sharding: all-reduced gradients == mean of the shards' gradients
    want = None
    elbo = []
    for shard in range(world):  # single-process reference: both shards, same eps
        bf.load_rng_state(bm, snap)
        for p in params:
            p.grad = None
        sl = slice(shard * B // world, (shard + 1) * B // world)
        loss, lq, lp = local_step(bm, ids_all[sl], labels_all[sl], S)
        loss.backward()
        live = [p for p in params if p.grad is not None]  # MOPED priors are Parameters that never get a gradient
        grads = [p.grad.detach().float().clone() for p in live]
        want = grads if want is None else [a + b for a, b in zip(want, grads)]
        elbo.append((lq, lp))
    want = [w / world for w in want]
    assert torch.equal(elbo[0][0], elbo[1][0]) and torch.equal(elbo[0][1], elbo[1][1])  # functions of the weights only
    bf.load_rng_state(bm, snap)
    for p in params:
        p.grad = None
    sync = parallel.GradSync(bm)
    sl = slice(rank * B // world, (rank + 1) * B // world)
    loss, lq, lp = local_step(bm, ids_all[sl], labels_all[sl], S)
    loss.backward()
    odd = [n for n, p in bm.named_parameters() if p.requires_grad and p.grad is None and any(p is q for q in sync._order)]
    if odd and rank == 0:
        print("tensors whose hook fired without a gradient:", odd[:8], flush=True)
    sync.finish()
    sync.remove()
    both = [torch.empty_like(lq) for _ in range(world)]
    dist.all_gather(both, lq)
    assert torch.equal(both[0], both[1]), "log q differs between ranks under batch sharding"
    assert [id(p) for p in params if p.grad is not None] == [id(p) for p in live]
    worst = max(rel(p.grad.float(), w) for p, w in zip(live, want))
    assert worst < 2e-3, f"batch-sharded all-reduce: worst relative gradient error {worst}"
    assert 0 < sync.bytes_last_step <= sum(p.numel() * p.grad.element_size() for p in live) + 64 * len(live)

    # ---- 3. sample sharding: S samples split over the ranks == single-process S-sample step
    bf.manual_seed(5000)
    S_local = parallel.shard_samples(S)
    assert S_local == S // world
    seeds = [None] * world
    dist.all_gather_object(seeds, runtime.seed())
    assert seeds[0] != seeds[1]
    snap = bf.rng_state(bm)
    for p in params:
        p.grad = None
    raws, kls, lqs, lps = [], [], [], []
    for r in range(world):  # emulation: this process plays both ranks, one backward through both forwards
        bf.load_rng_state(bm, dict(snap, seed=seeds[r]))
        with bf.mc_samples(S_local):
            logits = bm(input_ids=ids_all.repeat(S_local, 1)).logits
        raws.append(logits.float().view(S_local, B, -1))
        lq, lp = bm.log_variational_posterior(), bm.log_prior()
        kls.append((lq.sum() - lp.sum()) / S)
        lqs.append(lq.detach().sum()), lps.append(lp.detach().sum())
    mean = torch.cat(raws).mean(0)
    loss_ref = sum(kls) / N_BATCHES + torch.nn.functional.cross_entropy(mean, labels_all)
    loss_ref.backward()
    want = [p.grad.detach().float().clone() for p in live]
    lq_ref, lp_ref = sum(lqs), sum(lps)
    # the distributed run
    bf.load_rng_state(bm, dict(snap, seed=seeds[rank]))
    for p in params:
        p.grad = None
    sync = parallel.GradSync(bm, average=False)
    loss, lq, lp = local_step(bm, ids_all, labels_all, S_local, S_total=S, group_mean=True)
    loss.backward()
    sync.finish()
    sync.remove()
    lq_all, lp_all = parallel.all_reduce_elbo(lq, lp)
    assert abs(float(lq_all) - float(lq_ref)) <= 1e-5 * abs(float(lq_ref))
    assert abs(float(lp_all) - float(lp_ref)) <= 1e-5 * abs(float(lp_ref))
    worst = max(rel(p.grad.float(), w) for p, w in zip(live, want))
    assert worst < 2e-3, f"sample-sharded step: worst relative gradient error {worst}"
    # loss: CE part identical on every rank; the KL part is this rank's share
    kl_local = (lq.sum() - lp.sum()) / S / N_BATCHES
    kl_total = (lq_all - lp_all) / S / N_BATCHES
    assert abs(float(loss - kl_local + kl_total) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))

    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("MULTIRANK OK: identical weights, batch-sharded all-reduce == mean of shards, sample-sharded step == "
              "single-process S-sample step", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
