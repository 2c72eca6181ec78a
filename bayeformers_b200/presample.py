"""Multi-tensor sampling: every variational tensor of a model -- Linear weights and
biases, LayerNorm gamma / beta, Embedding tables -- drawn by ONE launch of
`bf_sample_kl_fwd_multi` per forward, instead of
two launches per layer.  Embedding tables take part with "no output": their log q /
log p are reduced over the whole table here, their rows are sampled on lookup
(`bf_embedding_fwd`) from the same stream.

    bm = to_bayesian(model, ...).to("cuda")
    bf.enable_presample(bm)        # opt-in; results are those of the per-layer path
    with bf.mc_samples(S): out = bm(**inputs)

`bnn.Model.forward` then samples all weights / biases and reduces all log q /
log p before the host model runs; each Bayesian layer's forward picks up its slice.
The eps stream is the usual Philox stream with step = 0x80000000 | run index, so
draws never collide with those of the per-layer path; backward regenerates eps
from the (seed, tensor_id, step) recorded here.  Layers with an injected eps
source (parity tests) or unsupported priors make the whole model fall back to the
per-layer path for that forward.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch
from torch.distributions.normal import Normal

from . import _lib, ops, runtime
from ._lib import BF_BF16, BF_F32, BF_PRIOR_GAUSSIAN, BfTensorDesc


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class Presampler:
    def __init__(self, model: torch.nn.Module) -> None:
        from .nn.layers.common import BayesianLayer

        self.layers: List[torch.nn.Module] = [m for m in model.modules() if isinstance(m, BayesianLayer)]
        self._sig = None
        self._runs = 0

    # ---- static tables (rebuilt when a pointer, S or a dtype changes) ---------------------
    def _tensors(self, S: int):
        from .nn.layers.embedding import Embedding
        from .nn.layers.linear import Linear
        from .nn.parameters.gaussian import Gaussian, prior_spec_of

        out = []  # (layer index, gaussian, prior spec, w dtype; None = log-probs only)
        for li, layer in enumerate(self.layers):
            if isinstance(layer, Linear):
                N, K = layer.weight.mu.shape
                tc = layer._gemm_dtype() == torch.bfloat16 and ops.tc_eligible(N, K)
                wdt = torch.bfloat16 if tc else torch.float32
            elif isinstance(layer, Embedding):
                wdt = None
            else:  # LayerNorm: fp32 affine
                wdt = torch.float32
            out.append((li, layer.weight, prior_spec_of(layer.weight_prior), wdt))
            if isinstance(getattr(layer, "bias", None), Gaussian):
                out.append((li, layer.bias, prior_spec_of(layer.bias_prior), torch.float32))
        return out

    def _signature(self, S: int, tensors):
        sig = [S]
        for li, g, pr, dt in tensors:
            sc = g.sigma_cache()
            sig += [g.mu.data_ptr(), g.rho.data_ptr(), pr.kind, 0 if pr.mu is None else pr.mu.data_ptr(),
                    0 if pr.rho is None else pr.rho.data_ptr(), pr.sigma1, dt, 0 if sc is None else sc.data_ptr()]
        return tuple(sig)

    def _build(self, S: int, tensors, dev: torch.device) -> None:
        lib = _lib.load()
        cq = lib.bf_sample_kl_multi_chunk_quads()
        descs = (BfTensorDesc * len(tensors))()
        chunks: List[int] = []
        slot_ranges: List[int] = []
        offsets = []
        off = 0
        cur_slot, slot_begin = None, 0
        for ti, (li, g, pr, dt) in enumerate(tensors):
            if li != cur_slot:
                if cur_slot is not None:
                    slot_ranges += [slot_begin, len(chunks) // 2]
                cur_slot, slot_begin = li, len(chunks) // 2
            n = g.mu.numel()
            esz = 2 if dt == torch.bfloat16 else 4
            d = descs[ti]
            d.mu, d.rho = g.mu.data_ptr(), g.rho.data_ptr()
            d.prior_mu = pr.mu.data_ptr() if (pr.kind == BF_PRIOR_GAUSSIAN and pr.mu is not None) else None
            d.prior_rho = pr.rho.data_ptr() if (pr.kind == BF_PRIOR_GAUSSIAN and pr.rho is not None) else None
            # byte offset into the per-run arena (w_base argument); all-ones = log-probs only (Embedding tables)
            d.w_out = off if dt is not None else ctypes.c_void_p(-1).value
            d.n, d.w_stride = n, n
            d.tensor_id, d.step = g.tensor_id, 0
            d.prior_kind, d.w_dtype = pr.kind, (BF_BF16 if dt == torch.bfloat16 else BF_F32)
            d.pi, d.sigma1, d.sigma2 = pr.pi, pr.sigma1, pr.sigma2
            sc = g.sigma_cache()  # softplus(rho) kept current by bf.optim.ClipAdamW: read instead of rho
            d.sigma = None if sc is None else sc.data_ptr()
            ptrs = [d.mu, d.rho, d.prior_mu or 0, d.prior_rho or 0, d.sigma or 0]
            d.vec = int(n % 4 == 0 and (n * esz) % 16 == 0 and all(p % 16 == 0 for p in ptrs))
            offsets.append((off, n, dt))
            if dt is not None:
                off += _align(S * n * esz)
            nquad = max((n + 3) // 4, 1)
            for q0 in range(0, nquad, cq):
                chunks += [ti, q0]
        slot_ranges += [slot_begin, len(chunks) // 2]
        self.arena_bytes = max(off, 256)
        self.offsets = offsets
        self.tensors = tensors
        self.n_chunks, self.n_slots = len(chunks) // 2, len(slot_ranges) // 2
        self.prior_mask = 0
        for _, _, pr, _ in tensors:
            self.prior_mask |= 1 << pr.kind
        raw = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8)
        self.d_descs = raw.to(dev)
        self.d_chunks = torch.tensor(chunks, dtype=torch.int32).to(dev)
        self.d_slots = torch.tensor(slot_ranges, dtype=torch.int32).to(dev)
        self.d_ws = torch.empty(lib.bf_sample_kl_multi_workspace_bytes(self.n_chunks), dtype=torch.uint8, device=dev)

    # ---- per forward ---------------------------------------------------------------------
    def run(self, S: int) -> bool:
        """Sample everything for this forward; False (and nothing done) when the model must use
        the per-layer path."""
        from .nn.parameters.gaussian import Gaussian

        if not self.layers:
            return False
        for layer in self.layers:
            for g in (layer.weight, getattr(layer, "bias", None)):
                if isinstance(g, Gaussian) and not isinstance(g.normal, Normal):
                    return False  # injected eps: per-layer path
        dev = self.layers[0].weight.mu.device
        if dev.type != "cuda":
            return False
        tensors = self._tensors(S)
        sig = self._signature(S, tensors)
        if sig != self._sig:
            self._build(S, tensors, dev)
            self._sig = sig
        lib = _lib.load()
        self._runs += 1
        step = 0x80000000 | (self._runs & 0x7FFFFFFF)
        arena = torch.empty(self.arena_bytes, dtype=torch.uint8, device=dev)  # fresh: earlier draws may still be
        logq = torch.empty((self.n_slots, S), dtype=torch.float32, device=dev)  # alive in autograd graphs
        logp = torch.empty((self.n_slots, S), dtype=torch.float32, device=dev)
        seed = runtime.seed()
        # algorithmic bytes: mu + rho (+ prior mu when it is a separate array, + prior rho when not constant) read,
        # S samples written
        nbytes = 0.0
        for (_, g, pr, dt), (_, n, _) in zip(self.tensors, self.offsets):
            p_bytes = 0
            if pr.kind == BF_PRIOR_GAUSSIAN:
                p_bytes = (4 if (pr.mu is not None and pr.mu.data_ptr() != g.mu.data_ptr()) else 0) + \
                          (4 if pr.rho is not None else 0)
            nbytes += n * (8 + p_bytes + (0 if dt is None else S * (2 if dt == torch.bfloat16 else 4)))
        rc = ops._timed("sample_kl_fwd", nbytes, dev, lambda: lib.bf_sample_kl_fwd_multi(
            self.d_descs.data_ptr(), self.d_chunks.data_ptr(), self.n_chunks, self.d_slots.data_ptr(), self.n_slots,
            self.prior_mask, S, seed, step, logq.data_ptr(), logp.data_ptr(), self.d_ws.data_ptr(), arena.data_ptr(),
            ops._stream(dev)))
        _lib.check(rc, "bf_sample_kl_fwd_multi")
        n_sc = bin(S).count("1") if S <= 8 else (S // 8 + bin(S % 8).count("1"))
        ops.stats["launches"] += (bin(self.prior_mask).count("1") + 1) * n_sc
        # hand every layer its slice
        means = (None, None)
        if S > 1:  # 0-dim values of the layers' registered scalars (see BayesianLayer._publish), all layers at once
            means = (logq.mean(1), logp.mean(1))
        per_layer = {}
        for (li, g, pr, dt), (off, n, _) in zip(self.tensors, self.offsets):
            w = None
            if dt is not None:
                esz = 2 if dt == torch.bfloat16 else 4
                w = arena[off:off + S * n * esz].view(dt).view((S,) + tuple(g.mu.shape))
            per_layer.setdefault(li, []).append((w, ops.StreamSpec(seed=seed, tensor_id=g.tensor_id, step=step)))
        for li, layer in enumerate(self.layers):
            items = per_layer[li]
            (W, w_stream) = items[0]
            (b, b_stream) = items[1] if len(items) > 1 else (None, ops.StreamSpec())
            m = None if means[0] is None else (means[0][li], means[1][li])
            layer._presampled = (S, W, b, logq[li], logp[li], w_stream, b_stream, m)
        return True


def enable_presample(model: torch.nn.Module, flag: bool = True) -> torch.nn.Module:
    """Switch multi-tensor sampling on/off for a `bnn.Model` (see module docstring)."""
    from .nn.model import Model

    if not isinstance(model, Model):
        raise TypeError("enable_presample expects the bnn.Model returned by to_bayesian")
    model._presampler = Presampler(model) if flag else None
    return model
