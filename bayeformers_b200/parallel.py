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
    # same eps on every rank, but independent dropout masks on each rank's own rows (like per-rank torch generators)
    runtime.set_dropout_salt(dist.get_rank(group))
    return runtime.seed()


class _Bucket:
    __slots__ = ("flat", "params", "pending", "handle")

    def __init__(self, flat, params):
        self.flat, self.params, self.pending, self.handle = flat, params, len(params), None


class GradSync:
    """Gradient reduction for a model containing Bayesian layers.

        sync = GradSync(model)            # after to_bayesian(...).to(device)
        sync.zero_grad()                  # instead of optimizer.zero_grad() (bucketed mode keeps .grad alive)
        loss.backward()
        sync.finish()                     # before clip / optimizer.step()

    bucketed (default when world_size > 1): after the FIRST backward the gradients of the tensors that actually received
    one (MOPED priors are trainable-looking Parameters that never do, quirk Q5) move into at most `buckets` flat
    buffers per dtype; from then on every such `p.grad` is a persistent view into its bucket, autograd accumulates in
    place, `zero_grad()` is one memset per bucket and `finish()` all-reduces a handful of large messages -- instead of one
    NCCL launch per tensor (74 for BERT-base).
    overlap=False (default): the buckets are reduced in `finish()`, after backward.  The whole message of a BERT-base
    step (~390 MB) is about 1 ms over NVLink 5 / NVSwitch, while a collective launched DURING backward shares the SMs
    with persistent one-CTA-per-SM contractions: the CTAs of a contraction that cannot be scheduled next to NCCL's
    channel CTAs start late, and the contraction ends late by as much as the collective lasts (measured: +8..11 ms per
    step for the overlapped variants).  overlap=True launches each bucket's all-reduce from the post-accumulate hook of
    its last tensor.
    bucketed=False (the round-1 behaviour): tensors with at least `large_numel` elements are all-reduced in place from a
    post-accumulate-grad hook, smaller ones flattened into a single message in `finish()`.
    `average=False` sums instead (sample sharding).  With world_size == 1 everything is a no-op.
    Gradient accumulation over several backward passes per step is not supported (each `finish()` reduces what is there).
    """

    def __init__(self, model: torch.nn.Module, group=None, large_numel: int = 1 << 18, average: bool = True,
                 bucketed: Optional[bool] = None, buckets: int = 4, overlap: bool = False):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.large_numel = int(large_numel)
        self.average = average
        self.overlap = bool(overlap)
        self.n_buckets = int(buckets)
        self.params: List[torch.nn.Parameter] = [p for p in model.parameters() if p.requires_grad]
        self._handles = []
        self._small: List[torch.nn.Parameter] = []
        self._hooks = []
        self.bytes_last_step = 0
        self._bytes = 0
        self.bucketed = (self.world > 1) if bucketed is None else (bool(bucketed) and self.world > 1)
        self.buckets: List[_Bucket] = []
        self._bucket_of = {}
        self._order: List[torch.nn.Parameter] = []  # first backward: the order gradients appear in
        if self.world > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(
                    self._on_grad_bucketed if self.bucketed else self._on_grad))

    # NCCL has a native AVG; gloo does not -> SUM then scale
    def _op(self):
        if self.average and dist.get_backend(self.group) == "nccl":
            return dist.ReduceOp.AVG, False
        return dist.ReduceOp.SUM, self.average

    # ---- bucketed mode ---------------------------------------------------------------------
    def _make_buckets(self) -> None:
        """Called by the first finish(): `self._order` lists the tensors that received a gradient, in the order backward
        produced them.  Their gradients move into flat buffers (values kept)."""
        by_dtype = {}
        seen = set()
        for p in self._order:
            if id(p) in seen or p.grad is None:  # (a hook may fire for a tensor whose gradient was dropped again)
                continue
            seen.add(id(p))
            by_dtype.setdefault((p.grad.dtype, p.grad.device), []).append(p)
        for (dtype, dev), plist in by_dtype.items():
            total = sum(p.numel() for p in plist)
            target = max((total + self.n_buckets - 1) // self.n_buckets, 1)
            groups, cur, cur_n = [], [], 0
            for p in plist:
                cur.append(p)
                cur_n += p.numel()
                if cur_n >= target:
                    groups.append(cur)
                    cur, cur_n = [], 0
            if cur:
                groups.append(cur)
            for g in groups:
                # every view starts 16 B aligned (vector loads of the optimizer / norm kernels)
                offs, n = [], 0
                for p in g:
                    offs.append(n)
                    n += (p.numel() + 7) // 8 * 8
                flat = torch.zeros(n, dtype=dtype, device=dev)
                for p, o in zip(g, offs):
                    view = flat[o:o + p.numel()].view_as(p)
                    view.copy_(p.grad)
                    p.grad = view
                b = _Bucket(flat, g)
                b.pending = 0  # this step's gradients are complete
                self.buckets.append(b)
                for p in g:
                    self._bucket_of[id(p)] = b

    def _on_grad_bucketed(self, p: torch.nn.Parameter) -> None:
        b = self._bucket_of.get(id(p))
        if b is None:
            if not self.buckets:
                self._order.append(p)  # first backward: remember who gets gradients, and in which order
            return
        b.pending -= 1
        if b.pending == 0 and self.overlap:
            op, scale_after = self._op()
            b.handle = (dist.all_reduce(b.flat, op=op, group=self.group, async_op=True), scale_after)
            self._bytes += b.flat.numel() * b.flat.element_size()

    def zero_grad(self) -> None:
        """Bucketed mode: one memset per bucket (the .grad views stay alive), tensors outside the buckets get None.
        Otherwise: set every .grad to None."""
        if self.bucketed and self.buckets:
            for b in self.buckets:
                b.flat.zero_()
            for p in self.params:
                if id(p) not in self._bucket_of:
                    p.grad = None
        else:
            for p in self.params:
                p.grad = None

    # ---- per-tensor mode -------------------------------------------------------------------
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
        """Reduce (or wait for) this step's gradients."""
        if self.world == 1:
            return
        if self.bucketed:
            if not self.buckets:
                self._make_buckets()
            op, scale_after = self._op()
            nvtx = self.buckets and self.buckets[0].flat.is_cuda
            if nvtx:
                torch.cuda.nvtx.range_push("bf:grad_allreduce")
            for b in self.buckets:
                if b.handle is None:
                    b.handle = (dist.all_reduce(b.flat, op=op, group=self.group, async_op=True), scale_after)
                    self._bytes += b.flat.numel() * b.flat.element_size()
            # a tensor that received its first gradient after the buckets were built (rare): reduced on its own
            late = [p for p in self.params if p.grad is not None and id(p) not in self._bucket_of]
            for p in late:
                dist.all_reduce(p.grad, op=op, group=self.group)
                if scale_after:
                    p.grad.div_(self.world)
                self._bytes += p.grad.numel() * p.grad.element_size()
            for b in self.buckets:
                h, sa = b.handle
                h.wait()
                if sa:
                    b.flat.div_(self.world)
                b.handle, b.pending = None, len(b.params)
            if nvtx:
                torch.cuda.nvtx.range_pop()
            self.bytes_last_step, self._bytes = self._bytes, 0
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

    def allreduce_alone_ms(self, iters: int = 5) -> Optional[float]:
        """Device time of all-reducing one step's gradient message with nothing else running (CUDA events), for the
        bench report.  Bucketed mode only; leaves the gradients scaled by world**iters when summing -- call it after
        the timed region."""
        if not (self.bucketed and self.world > 1 and self.buckets and self.buckets[0].flat.is_cuda):
            return None
        op, _ = self._op()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            for b in self.buckets:
                dist.all_reduce(b.flat, op=op, group=self.group)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

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
