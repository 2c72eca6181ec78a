"""Data-parallel plumbing for the variational layers: one process per GPU,
`torch.distributed` (NCCL over NVLink 5 / NVSwitch on the box, gloo in the CPU
tests).

What is and is not communicated:
  * weights and eps: NEVER -- every rank regenerates the same Philox stream
    from (seed, tensor_id, step, sample), so replicas draw bit-identical W;
  * log q / log p: functions of the weights only, hence already identical on
    every rank under batch sharding -- no reduction;
  * gradients of rho (and of mu / frequentist tensors when trainable): summed
    with all-reduce, launched per tensor as soon as autograd has produced it so
    the transfer overlaps the rest of backward.  The kernels hand autograd
    freshly written gradient buffers, which are reduced in place (no bucket
    copy for the large tensors); small tensors are coalesced into one flat
    message at the end.

The reference's only multi-GPU code is `torch.nn.DataParallel` in one example
(/root/reference/examples/bert_squad.py:245), which re-broadcasts all
parameters every forward and lets each replica draw different eps.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from . import runtime


def broadcast_seed(src: int = 0, group=None) -> int:
    """Make every rank use rank `src`'s eps seed (call once after init_process_group)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([runtime.seed() & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=dev)
    dist.broadcast(t, src=src, group=group)
    runtime.manual_seed(int(t.item()))
    return runtime.seed()


class GradSync:
    """Overlapped gradient averaging for a model containing Bayesian layers.

        sync = GradSync(model)            # after to_bayesian(...).to(device)
        loss.backward()
        sync.finish()                     # before clip / optimizer.step()

    Tensors with at least `large_numel` elements are all-reduced in place,
    asynchronously, from a post-accumulate-grad hook (overlaps with the
    remaining backward); smaller ones are flattened into a single message in
    `finish()`.  With world_size == 1 everything is a no-op.
    """

    def __init__(self, model: torch.nn.Module, group=None, large_numel: int = 1 << 18, average: bool = True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.large_numel = int(large_numel)
        self.average = average
        self.params: List[torch.nn.Parameter] = [p for p in model.parameters() if p.requires_grad]
        self._handles = []
        self._small: List[torch.nn.Parameter] = []
        self._hooks = []
        self.bytes_last_step = 0
        self._bytes = 0
        if self.world > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # NCCL has a native AVG; gloo does not -> SUM then scale
    def _op(self):
        if self.average and dist.get_backend(self.group) == "nccl":
            return dist.ReduceOp.AVG, False
        return dist.ReduceOp.SUM, self.average

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        g = p.grad
        if g is None:
            return
        if g.numel() >= self.large_numel and g.is_contiguous():
            op, scale_after = self._op()
            h = dist.all_reduce(g, op=op, group=self.group, async_op=True)
            self._handles.append((h, g if scale_after else None))
            self._bytes += g.numel() * g.element_size()
        else:
            self._small.append(p)

    def finish(self) -> None:
        """Wait for the in-flight reductions and reduce the coalesced small tensors."""
        if self.world == 1:
            return
        if self._small:
            grads = [p.grad for p in self._small]
            flat = torch._utils._flatten_dense_tensors(grads)
            op, scale_after = self._op()
            dist.all_reduce(flat, op=op, group=self.group)
            if scale_after:
                flat.div_(self.world)
            for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                g.copy_(f)
            self._bytes += flat.numel() * flat.element_size()
            self._small = []
        for h, g in self._handles:
            h.wait()
            if g is not None:
                g.div_(self.world)
        self._handles = []
        self.bytes_last_step, self._bytes = self._bytes, 0

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


# ---- sample sharding (second mode: S >= G, e.g. S = 8 or 16 over 8 GPUs) -------------------
def shard_samples(S: int, group=None) -> int:
    """Give this rank its share of the S Monte-Carlo samples: returns S // world and re-keys the eps stream with a
    rank-distinct seed derived from rank 0's, so the ranks draw independent weight samples (under batch sharding
    they must draw IDENTICAL ones -- do not mix the two modes in one step).  Every rank then runs the full batch
    with `mc_samples(S // world)`; gradients are SUMMED (`GradSync(..., average=False)`), the ELBO scalars go through
    `all_reduce_elbo`, and the per-sample logits through `mean_over_samples` before the loss."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if S % world != 0:
        raise ValueError(f"mc_samples={S} is not a multiple of the world size {world}")
    if world > 1:
        base = broadcast_seed(0, group)
        rank = dist.get_rank(group)
        runtime.manual_seed((base + 0x9E3779B97F4A7C15 * (rank + 1)) & 0xFFFFFFFFFFFFFFFF)
    return S // world


class _MeanOverRanks(torch.autograd.Function):
    @staticmethod
    def forward(ctx, local_sum: torch.Tensor, total: int, group):
        out = local_sum.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        ctx.total = total
        return out / total

    @staticmethod
    def backward(ctx, g):
        # the loss is replicated on every rank and a function of the same mean: each rank only needs the
        # derivative with respect to ITS samples, g / S -- no communication in backward
        return g / ctx.total, None, None


def mean_over_samples(raw_local: torch.Tensor, S_total: int, group=None) -> torch.Tensor:
    """raw_local [S_local, B, ...] -> mean over all S_total samples held by all ranks, differentiable.  The
    reference's loss is the cross entropy of the MEAN over samples of the logits (examples/bert_glue.py:69,234), a
    non-linear function of that mean, so the logits are averaged across ranks BEFORE the loss (S*B*C floats)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return raw_local.sum(0) / S_total
    return _MeanOverRanks.apply(raw_local.sum(0), int(S_total), group)


def all_reduce_elbo(log_q: torch.Tensor, log_p: torch.Tensor, group=None):
    """Sample-sharded runs only (each rank owns different MC samples): sum the
    per-rank [log q, log p] pair.  Under batch sharding the scalars are already
    identical on every rank and this must NOT be called."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return log_q, log_p
    v = torch.stack([log_q.sum(), log_p.sum()])
    dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return v[0], v[1]
