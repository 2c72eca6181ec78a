"""torch.autograd.Function wrappers that call the C-ABI kernels.

PyTorch is plumbing here: it owns the device buffers and the stream; every
arithmetic step of the hot path runs in libbayeformers_b200.so.  There is no
eager/CPU substitute -- non-CUDA tensors raise.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import BF_BF16, BF_F32, BF_PRIOR_GAUSSIAN, BF_PRIOR_MIXTURE, BF_PRIOR_NONE

_workspaces: Dict[Tuple, torch.Tensor] = {}

# ---- instrumentation (bench.py): kernel-launch counter and optional per-call CUDA-event timing
stats = {"launches": 0, "dgrad_accumulated": 0}
_timing = {"on": False, "records": []}


def enable_kernel_timing(flag: bool) -> None:
    """Bracket every contraction / sample+KL call with CUDA events recorded on
    the launching stream; `kernel_timing_summary()` turns them into durations."""
    _timing["on"] = bool(flag)
    _timing["records"] = []


_nvtx = {"on": False}


def enable_nvtx(flag: bool = True) -> None:
    """Wrap every C-ABI launch in an NVTX range named after its kernel family (sample_kl_fwd, gemm_fwd_tc,
    gemm_dgrad_tc, gemm_wgrad_fused_tc, resln_*, clip_adamw, ...): the ranges ncu / nsys group launches by."""
    _nvtx["on"] = bool(flag)


def _timed(name: str, work: float, dev, fn):
    """Run fn() (which launches on the current stream), counting and optionally timing it."""
    if _nvtx["on"]:
        torch.cuda.nvtx.range_push("bf:" + name)
        try:
            return _timed_inner(name, work, dev, fn)
        finally:
            torch.cuda.nvtx.range_pop()
    return _timed_inner(name, work, dev, fn)


def _timed_inner(name: str, work: float, dev, fn):
    if not _timing["on"]:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.current_stream(dev)
    e0.record(stream)
    out = fn()
    e1.record(stream)
    _timing["records"].append((name, work, e0, e1))
    return out


def kernel_timing_summary():
    """{name: {calls, ms, work}} -- call after torch.cuda.synchronize()."""
    out = {}
    for name, work, e0, e1 in _timing["records"]:
        d = out.setdefault(name, {"calls": 0, "ms": 0.0, "work": 0.0})
        d["calls"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["work"] += work
    return out


def _fwd_launches(n: int, S: int, aligned: bool) -> int:
    k, s = 0, S
    if aligned and n >= 4:
        while s > 0:  # chunks of 8/4/2/1 samples
            c = 8 if s >= 8 else (4 if s >= 4 else (2 if s >= 2 else 1))
            s -= c
            k += 1
    if not aligned or n % 4 != 0 or n == 0:
        k += S
    return k


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(dev: torch.device):
    return torch.cuda.current_stream(dev).cuda_stream


class on_device:
    """Make `dev` the current CUDA device for the duration of a group of C-ABI calls (the library launches on the
    current device: streams, SM counts and the step counter are per device).  A no-op -- one integer compare --
    in the usual one-process-per-GPU set-up where it already is."""
    __slots__ = ("idx", "prev")

    def __init__(self, dev: torch.device) -> None:
        self.idx = dev.index if dev.index is not None else torch.cuda.current_device()
        self.prev = -1

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False


def _guarded(fn):
    """Decorator for autograd.Function forward / backward bodies: run under `on_device` of the first CUDA tensor."""
    import functools

    @functools.wraps(fn)
    def wrapper(ctx, *args, **kwargs):
        dev = next((a.device for a in args if torch.is_tensor(a) and a.is_cuda), None)
        if dev is None:
            saved = getattr(ctx, "saved_tensors", ())
            dev = next((t.device for t in saved if torch.is_tensor(t) and t.is_cuda), None)
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(ctx, *args, **kwargs)
        with on_device(dev):
            return fn(ctx, *args, **kwargs)

    return wrapper


def _dt(t: torch.dtype) -> int:
    if t == torch.float32:
        return BF_F32
    if t == torch.bfloat16:
        return BF_BF16
    raise TypeError(f"unsupported dtype {t}")


def _aligned(t: torch.Tensor) -> torch.Tensor:
    """Dense tensor whose storage offset breaks 16-byte alignment (a slice of a larger buffer) -> private copy:
    the row kernels use 16-byte vector and bulk copies."""
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone()


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"bayeformers_b200: {what} is on {t.device}; the variational layers run on CUDA (sm_100a) only "
            "-- there is no CPU path. Move the model and inputs to a B200 device.")


def _workspace(kind: str, dev: torch.device, nbytes: int) -> torch.Tensor:
    """Zero-filled-once scratch, one per (kind, device, stream)."""
    key = (kind, dev.index, _stream(dev))
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(int(nbytes), 1024), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws


@dataclass
class PriorSpec:
    """What the kernel needs to evaluate log p(w) (and its derivative)."""
    kind: int = BF_PRIOR_NONE
    pi: float = 0.5
    sigma1: float = 1.0
    sigma2: float = 1.0
    mu: Optional[torch.Tensor] = None   # BF_PRIOR_GAUSSIAN
    rho: Optional[torch.Tensor] = None


@dataclass
class StreamSpec:
    """Identity of one eps draw: (seed, tensor_id, step); `eps` overrides the
    Philox stream with injected values of shape [S, *param_shape] (parity)."""
    seed: int = 0
    tensor_id: int = 0
    step: int = 0
    eps: Optional[torch.Tensor] = None


def _eps_arg(stream: StreamSpec, S: int, n: int, dev) -> Optional[torch.Tensor]:
    if stream.eps is None:
        return None
    e = stream.eps.to(device=dev, dtype=torch.float32).contiguous()
    if e.numel() != S * n:
        raise ValueError(f"injected eps has {e.numel()} elements, expected S*n = {S}*{n}")
    return e


def sample_kl_forward(mu, rho, prior: PriorSpec, stream: StreamSpec, S: int, w_dtype, logq, logp, accumulate: bool,
                      want_w: bool = True):
    """w[S, *mu.shape] and logq/logp[S] (+= when accumulate) in one kernel pass."""
    lib = _lib.load()
    dev = mu.device
    n = mu.numel()
    eps = _eps_arg(stream, S, n, dev)
    w = torch.empty((S,) + tuple(mu.shape), dtype=w_dtype, device=dev) if want_w else None
    ws = _workspace("sample_kl", dev, lib.bf_sample_kl_workspace_bytes(n, S))
    pmu = prior.mu if prior.kind == BF_PRIOR_GAUSSIAN else None
    prho = prior.rho if prior.kind == BF_PRIOR_GAUSSIAN else None
    nbytes = n * (8 + (0 if pmu is None else 4) + (0 if prho is None else 4) +
                  (S * (2 if w_dtype == torch.bfloat16 else 4) if want_w else 0)) + 8 * S
    rc = _timed("sample_kl_fwd", float(nbytes), dev, lambda: lib.bf_sample_kl_fwd(
        _ptr(mu), _ptr(rho), prior.kind, _ptr(pmu), _ptr(prho), prior.pi, prior.sigma1,
        prior.sigma2, n, S, stream.seed, stream.step, stream.tensor_id, _ptr(eps), _ptr(w),
        _dt(w_dtype), n, _ptr(logq), _ptr(logp), int(accumulate), _ptr(ws), _stream(dev)))
    _lib.check(rc, "bf_sample_kl_fwd")
    stats["launches"] += _fwd_launches(n, S, mu.data_ptr() % 16 == 0 and rho.data_ptr() % 16 == 0)
    return w


def sample_kl_backward(grad_w, mu, rho, prior: PriorSpec, stream: StreamSpec, S: int, g_logq, g_logp, need_mu: bool):
    lib = _lib.load()
    dev = rho.device
    n = rho.numel()
    eps = _eps_arg(stream, S, n, dev)
    g_rho = torch.empty_like(rho, dtype=torch.float32)
    g_mu = torch.empty_like(rho, dtype=torch.float32) if need_mu else None
    if grad_w is None and g_logq is None and g_logp is None:
        g_rho.zero_()
        if g_mu is not None:
            g_mu.zero_()
        return g_mu, g_rho
    gw_dt = BF_F32
    if grad_w is not None:
        grad_w = grad_w.contiguous()
        gw_dt = _dt(grad_w.dtype)
    pmu = prior.mu if prior.kind == BF_PRIOR_GAUSSIAN else None
    prho = prior.rho if prior.kind == BF_PRIOR_GAUSSIAN else None
    nbytes = n * ((0 if grad_w is None else S * grad_w.element_size()) + 8 + (4 if need_mu else 0))
    rc = _timed("sample_kl_bwd", float(nbytes), dev, lambda: lib.bf_sample_kl_bwd(
        _ptr(grad_w), gw_dt, n, _ptr(mu), _ptr(rho), prior.kind, _ptr(pmu), _ptr(prho),
        prior.pi, prior.sigma1, prior.sigma2, _ptr(g_logq), _ptr(g_logp), n, S, stream.seed,
        stream.step, stream.tensor_id, _ptr(eps), _ptr(g_mu), _ptr(g_rho), 0, _stream(dev)))
    _lib.check(rc, "bf_sample_kl_bwd")
    stats["launches"] += 1
    return g_mu, g_rho


def _kl_upstream(g, S, dev):
    if g is None:
        return torch.zeros(S, dtype=torch.float32, device=dev)
    return g.to(torch.float32).reshape(S).contiguous()


class SampleKL(torch.autograd.Function):
    """(mu, rho) -> (w[S,...], log q[S], log p[S]) for one variational tensor.
    Replaces Gaussian.sample + the two log_prob calls of the reference
    (gaussian.py:90-116,160-171) with one fused pass; backward regenerates eps."""

    @staticmethod
    @_guarded
    def forward(ctx, mu, rho, prior_mu, prior_rho, prior: PriorSpec, stream: StreamSpec, S: int, w_dtype, kl_grad):
        _require_cuda(mu, "variational parameter")
        prior = PriorSpec(prior.kind, prior.pi, prior.sigma1, prior.sigma2, prior_mu, prior_rho)
        dev = mu.device
        logq = torch.empty(S, dtype=torch.float32, device=dev)
        logp = torch.empty(S, dtype=torch.float32, device=dev)
        w = sample_kl_forward(mu.detach(), rho.detach(), prior, stream, S, w_dtype, logq, logp, False)
        ctx.save_for_backward(mu, rho, prior_mu, prior_rho)
        ctx.meta = (prior, stream, S, kl_grad)
        if not kl_grad:
            ctx.mark_non_differentiable(logq, logp)
        return w, logq, logp

    @staticmethod
    @_guarded
    def backward(ctx, gw, glq, glp):
        mu, rho, prior_mu, prior_rho = ctx.saved_tensors
        prior, stream, S, kl_grad = ctx.meta
        prior = PriorSpec(prior.kind, prior.pi, prior.sigma1, prior.sigma2, prior_mu, prior_rho)
        dev = rho.device
        if kl_grad:
            glq, glp = _kl_upstream(glq, S, dev), _kl_upstream(glp, S, dev)
        else:
            glq = glp = None
        if gw is not None:
            gw = gw.reshape(S, -1)
            if gw.dtype not in (torch.float32, torch.bfloat16):
                gw = gw.float()
        g_mu, g_rho = sample_kl_backward(gw, mu.detach(), rho.detach(), prior, stream, S, glq, glp,
                                         ctx.needs_input_grad[0])
        return g_mu, g_rho, None, None, None, None, None, None, None


@dataclass
class PairSpec:
    """One (weight, bias) pair of small variational tensors sampled together (bnn.LayerNorm's gamma / beta)."""
    S: int
    kl_grad: bool
    w_prior: PriorSpec
    b_prior: PriorSpec
    w_stream: StreamSpec
    b_stream: StreamSpec = field(default_factory=StreamSpec)
    presampled: Optional[Tuple] = None  # (w[S,n] fp32, b[S,n] fp32 or None, logq[S], logp[S]) from the Presampler


class SamplePair(torch.autograd.Function):
    """(w_mu, w_rho, b_mu, b_rho) -> (w[S,...], b[S,...] or None, log q[S], log p[S]): weight then bias accumulated
    into one pair of scalars, as Linear does (linear.py:97-102).  Fresh draw or the multi-tensor sampler's; backward
    regenerates eps for both tensors from their recorded streams."""

    @staticmethod
    @_guarded
    def forward(ctx, w_mu, w_rho, b_mu, b_rho, wp_mu, wp_rho, bp_mu, bp_rho, spec: PairSpec):
        _require_cuda(w_mu, "variational parameter")
        S, dev = spec.S, w_mu.device
        has_bias = b_mu is not None
        if spec.presampled is not None:
            w, b, logq, logp = spec.presampled
            spec.presampled = None  # outputs of this node must not be reachable from its own ctx (see BayesLinear)
        else:
            logq = torch.empty(S, dtype=torch.float32, device=dev)
            logp = torch.empty(S, dtype=torch.float32, device=dev)
            wp = PriorSpec(spec.w_prior.kind, spec.w_prior.pi, spec.w_prior.sigma1, spec.w_prior.sigma2, wp_mu, wp_rho)
            w = sample_kl_forward(w_mu.detach(), w_rho.detach(), wp, spec.w_stream, S, torch.float32, logq, logp, False)
            b = None
            if has_bias:
                bp = PriorSpec(spec.b_prior.kind, spec.b_prior.pi, spec.b_prior.sigma1, spec.b_prior.sigma2, bp_mu,
                               bp_rho)
                b = sample_kl_forward(b_mu.detach(), b_rho.detach(), bp, spec.b_stream, S, torch.float32, logq, logp,
                                      True)
        ctx.save_for_backward(w_mu, w_rho, b_mu, b_rho, wp_mu, wp_rho, bp_mu, bp_rho)
        ctx.spec = spec
        if not spec.kl_grad:
            ctx.mark_non_differentiable(logq, logp)
        if not has_bias:
            b = torch.empty(0, device=dev)  # placeholder output (autograd outputs must be tensors)
            ctx.mark_non_differentiable(b)
        return w, b, logq, logp

    @staticmethod
    @_guarded
    def backward(ctx, gw, gb, glq, glp):
        w_mu, w_rho, b_mu, b_rho, wp_mu, wp_rho, bp_mu, bp_rho = ctx.saved_tensors
        spec = ctx.spec
        S, dev = spec.S, w_rho.device
        if spec.kl_grad:
            glq, glp = _kl_upstream(glq, S, dev), _kl_upstream(glp, S, dev)
        else:
            glq = glp = None

        def one(g, mu, rho, prior, pmu, prho, stream, need_mu):
            pr = PriorSpec(prior.kind, prior.pi, prior.sigma1, prior.sigma2, pmu, prho)
            if g is not None:
                g = g.reshape(S, -1)
                if g.dtype not in (torch.float32, torch.bfloat16):
                    g = g.float()
            return sample_kl_backward(g, mu.detach(), rho.detach(), pr, stream, S, glq, glp, need_mu)

        g_wmu, g_wrho = one(gw, w_mu, w_rho, spec.w_prior, wp_mu, wp_rho, spec.w_stream, ctx.needs_input_grad[0])
        g_bmu = g_brho = None
        if b_mu is not None:
            g_bmu, g_brho = one(gb, b_mu, b_rho, spec.b_prior, bp_mu, bp_rho, spec.b_stream, ctx.needs_input_grad[2])
        return g_wmu, g_wrho, g_bmu, g_brho, None, None, None, None, None


@dataclass
class EmbeddingSpec:
    S: int
    kl_grad: bool
    prior: PriorSpec
    stream: StreamSpec
    out_dtype: torch.dtype
    padding_idx: Optional[int] = None
    presampled: Optional[Tuple] = None  # (logq[S], logp[S]) of the whole table from the Presampler


class EmbeddingFn(torch.autograd.Function):
    """ids [S*B, ...] -> (rows [S*B, ..., H], log q[S], log p[S]) of a Bayesian embedding table.
    The log-probs are reduced over the WHOLE table (one sample+KL pass that writes no weights), while only the
    looked-up rows are sampled (`bf_embedding_fwd`); backward is row-sparse and deterministic (`bf_embedding_bwd`),
    plus the dense KL terms when kl_grad is on."""

    @staticmethod
    @_guarded
    def forward(ctx, ids, mu, rho, prior_mu, prior_rho, spec: EmbeddingSpec):
        _require_cuda(mu, "embedding table")
        lib = _lib.load()
        dev = mu.device
        S = spec.S
        V, H = mu.shape
        if ids.device != dev:
            raise RuntimeError(f"bayeformers_b200: ids are on {ids.device}, the table on {dev}")
        idc = ids.detach().to(torch.int64).contiguous()
        n_tok = idc.numel()
        if idc.dim() < 1 or idc.shape[0] % S != 0:
            raise ValueError(f"ids of shape {tuple(ids.shape)} cannot be split into mc_samples={S} groups of rows")
        tps = max(n_tok // S, 1)
        prior = PriorSpec(spec.prior.kind, spec.prior.pi, spec.prior.sigma1, spec.prior.sigma2, prior_mu, prior_rho)
        if spec.presampled is not None:
            logq, logp = spec.presampled
            spec.presampled = None  # outputs of this node must not be reachable from its own ctx (see BayesLinear)
        else:
            logq = torch.empty(S, dtype=torch.float32, device=dev)
            logp = torch.empty(S, dtype=torch.float32, device=dev)
            sample_kl_forward(mu.detach(), rho.detach(), prior, spec.stream, S, torch.float32, logq, logp, False,
                              want_w=False)
        eps = _eps_arg(spec.stream, S, V * H, dev)
        out = torch.empty(tuple(idc.shape) + (H,), dtype=spec.out_dtype, device=dev)
        nbytes = float(n_tok * H * (8 + out.element_size()))
        rc = _timed("embedding_fwd", nbytes, dev, lambda: lib.bf_embedding_fwd(
            _ptr(idc), n_tok, tps, _ptr(mu), _ptr(rho), V, H, spec.stream.seed, spec.stream.step, spec.stream.tensor_id,
            _ptr(eps), _ptr(out), _dt(spec.out_dtype), _stream(dev)))
        _lib.check(rc, "bf_embedding_fwd")
        stats["launches"] += 1
        ctx.save_for_backward(idc, mu, rho, prior_mu, prior_rho)
        ctx.meta = (spec, tps)
        ctx.mark_non_differentiable(*( [] if spec.kl_grad else [logq, logp]))
        return out, logq, logp

    @staticmethod
    @_guarded
    def backward(ctx, gout, glq, glp):
        lib = _lib.load()
        idc, mu, rho, prior_mu, prior_rho = ctx.saved_tensors
        spec, tps = ctx.meta
        S, dev = spec.S, mu.device
        V, H = mu.shape
        need_mu = ctx.needs_input_grad[1]
        prior = PriorSpec(spec.prior.kind, spec.prior.pi, spec.prior.sigma1, spec.prior.sigma2, prior_mu, prior_rho)
        if spec.kl_grad:
            # dense part: the KL terms touch every element of the table
            g_mu, g_rho = sample_kl_backward(None, mu.detach(), rho.detach(), prior, spec.stream, S,
                                             _kl_upstream(glq, S, dev), _kl_upstream(glp, S, dev), need_mu)
        else:
            g_rho = torch.zeros_like(rho, dtype=torch.float32)
            g_mu = torch.zeros_like(rho, dtype=torch.float32) if need_mu else None
        n_tok = idc.numel()
        if gout is not None and n_tok > 0:
            g = gout.reshape(n_tok, H)
            if g.dtype not in (torch.float32, torch.bfloat16):
                g = g.float()
            g = _aligned(g)
            sorted_ids, perm = torch.sort(idc.reshape(-1), stable=True)  # index bookkeeping; arithmetic is ours
            eps = _eps_arg(spec.stream, S, V * H, dev)
            ws = _workspace("embedding_bwd", dev, lib.bf_embedding_bwd_workspace_bytes(n_tok, H))
            pad = -1 if spec.padding_idx is None else int(spec.padding_idx) % V
            nbytes = float(n_tok * H * (g.element_size() + 8))
            rc = _timed("embedding_bwd", nbytes, dev, lambda: lib.bf_embedding_bwd(
                _ptr(g), _dt(g.dtype), _ptr(sorted_ids), _ptr(perm), n_tok, tps, _ptr(rho), V, H, pad, spec.stream.seed,
                spec.stream.step, spec.stream.tensor_id, _ptr(eps), _ptr(g_mu), _ptr(g_rho), _ptr(ws), _stream(dev)))
            _lib.check(rc, "bf_embedding_bwd")
            stats["launches"] += 2
        return None, g_mu, g_rho, None, None, None


def embedding_supported(H: int) -> bool:
    return bool(_lib.load().bf_embedding_supported(int(H)))


@dataclass
class LinearSpec:
    S: int
    gemm_dtype: torch.dtype
    kl_grad: bool
    w_prior: PriorSpec
    b_prior: PriorSpec
    w_stream: StreamSpec
    b_stream: StreamSpec = field(default_factory=StreamSpec)
    presampled: Optional[Tuple] = None  # (W[S,N,K], b[S,N] or None, logq[S], logp[S]) from presample.Presampler
    activation: Optional[str] = None    # None | "gelu": y = act(x w^T + b) with the activation fused where possible
    bias_grad_box: Optional[list] = None  # filled by ResidualLayerNormFn.backward with sum_m gy[s][m][:] ([S, N] fp32)
    sink: Optional[object] = None         # runtime.GradSink of the input: add dx into its buffer instead of returning it
    gelu_in: Optional[object] = None      # runtime.GeluLink of the INPUT (it is gelu(z) of a fused layer): dgrad applies gelu'
    gelu_out: Optional[list] = None       # filled by forward with the GeluLink of this layer's fused-GELU output


def split_bf16x2(t: torch.Tensor):
    """fp32 tensor -> bf16 [2, *shape]: out[0] = hi, out[1] = lo with hi + lo == t to 2^-17 relative (operands of the
    fp32x3 contractions)."""
    lib = _lib.load()
    t = t.contiguous()
    out = torch.empty((2,) + tuple(t.shape), dtype=torch.bfloat16, device=t.device)
    rc = lib.bf_split_bf16x2(_ptr(t), _ptr(out[0]), _ptr(out[1]), t.numel(), _stream(t.device))
    _lib.check(rc, "bf_split_bf16x2")
    stats["launches"] += 1
    return out


def tc_eligible(N: int, K: int) -> bool:
    """Shapes the TMA-fed tcgen05 path accepts (16-byte global row strides)."""
    return N % 8 == 0 and K % 8 == 0


class BayesLinear(torch.autograd.Function):
    """y, log q[S], log p[S] = BayesLinear(x, mu/rho of weight and bias, priors).

    Forward : fused sample+KL for W (and b) -> batched-over-S contraction.
    Backward: dgrad contraction; wgrad contraction whose epilogue regenerates
              eps and emits dmu / drho directly (bf16 mode), or fp32 wgrad +
              the stand-alone variational backward (fp32 parity mode).
    Replaces Linear.forward (linear.py:83-104) and its autograd graph."""

    @staticmethod
    @_guarded
    def forward(ctx, x, w_mu, w_rho, b_mu, b_rho, wp_mu, wp_rho, bp_mu, bp_rho, spec: LinearSpec):
        _require_cuda(x, "input")
        _require_cuda(w_mu, "weight")
        lib = _lib.load()
        dev = x.device
        S = spec.S
        N, K = w_mu.shape
        rows = x.numel() // K
        if x.shape[-1] != K or rows % S != 0:
            raise ValueError(f"input {tuple(x.shape)} incompatible with weight {(N, K)} and mc_samples={S}")
        M = rows // S
        use_tc = spec.gemm_dtype == torch.bfloat16 and tc_eligible(N, K)
        # "fp32x3": fp32 operands, split into bf16 (hi, lo) pairs, three tensor-core passes per tile (1e-5 parity mode)
        use_x3 = spec.gemm_dtype == "fp32x3" and tc_eligible(N, K) and M > 0
        cdt = torch.bfloat16 if use_tc else torch.float32
        xg = x.detach().reshape(S, M, K).to(cdt).contiguous()
        has_bias = b_mu is not None

        if spec.presampled is not None:
            W, b, logq, logp = spec.presampled
            # `spec` is kept by ctx for backward and logq / logp become OUTPUTS of this node: leaving them in the spec
            # would make the node reference itself (node -> spec -> output -> grad_fn = node), a cycle only the garbage
            # collector frees -- one sampled-weight arena plus the saved activations leaked per step
            spec.presampled = None
            if W.dtype != cdt or (has_bias and b is None):
                raise RuntimeError("presampled weights do not match this layer's GEMM mode")
        else:
            logq = torch.empty(S, dtype=torch.float32, device=dev)
            logp = torch.empty(S, dtype=torch.float32, device=dev)
            w_prior = PriorSpec(spec.w_prior.kind, spec.w_prior.pi, spec.w_prior.sigma1, spec.w_prior.sigma2, wp_mu,
                                wp_rho)
            W = sample_kl_forward(w_mu.detach(), w_rho.detach(), w_prior, spec.w_stream, S, cdt, logq, logp, False)
            b = None
            if has_bias:
                b_prior = PriorSpec(spec.b_prior.kind, spec.b_prior.pi, spec.b_prior.sigma1, spec.b_prior.sigma2,
                                    bp_mu, bp_rho)
                b = sample_kl_forward(b_mu.detach(), b_rho.detach(), b_prior, spec.b_stream, S, torch.float32, logq,
                                      logp, True)
        out_dtype = x.dtype if (use_tc and x.dtype in (torch.float32, torch.bfloat16)) else torch.float32
        act = spec.activation
        if act not in (None, "gelu"):
            raise ValueError(f"unsupported activation {act!r}")
        z = None  # pre-activation kept for backward when an activation is applied
        fused_act = (act == "gelu" and use_tc and has_bias and out_dtype == torch.bfloat16 and M > 0
                     and bool(lib.bf_linear_fwd_gelu_supported(S, M, N, K)))
        y = torch.empty((S, M, N), dtype=out_dtype, device=dev)
        if use_x3:
            # the (hi, lo) pairs replace the fp32 operands in what backward keeps (same bytes)
            xg, W = split_bf16x2(xg), split_bf16x2(W)
            rc = _timed("gemm_fwd_x3", 2.0 * S * M * N * K, dev, lambda: lib.bf_linear_fwd_x3(
                _ptr(xg[0]), _ptr(xg[1]), _ptr(W[0]), _ptr(W[1]), _ptr(b), _ptr(y), S, M, N, K, _stream(dev)))
            _lib.check(rc, "bf_linear_fwd_x3")
            stats["launches"] += 1
        elif fused_act:
            z = torch.empty((S, M, N), dtype=torch.bfloat16, device=dev)
            rc = _timed("gemm_fwd_tc", 2.0 * S * M * N * K, dev, lambda: lib.bf_linear_fwd_gelu(
                _ptr(xg), _ptr(W), _ptr(b), _ptr(z), _ptr(y), S, M, N, K, _stream(dev)))
            _lib.check(rc, "bf_linear_fwd_gelu")
            stats["launches"] += 1
        elif M > 0:
            rc = _timed("gemm_fwd_" + ("tc" if use_tc else "f32"), 2.0 * S * M * N * K, dev, lambda: lib.bf_linear_fwd(
                _ptr(xg), _ptr(W), _ptr(b), _ptr(y), S, M, N, K, _dt(cdt), _dt(out_dtype), _stream(dev)))
            _lib.check(rc, "bf_linear_fwd")
            stats["launches"] += 1
        if act == "gelu" and not fused_act:  # shapes the fused kernel does not take: same maths, separate pass
            z = y
            y = torch.nn.functional.gelu(z)
        ctx.save_for_backward(xg, W, w_mu, w_rho, b_mu, b_rho, wp_mu, wp_rho, bp_mu, bp_rho, z)
        link_out = None
        if fused_act and spec.gelu_out is not None:
            from .runtime import GeluLink
            link_out = GeluLink(z)
            spec.gelu_out.append(link_out)
        ctx.meta = (spec, use_tc, M, x.dtype, tuple(x.shape), fused_act, use_x3, link_out)
        if not spec.kl_grad:
            ctx.mark_non_differentiable(logq, logp)
        if x.dtype != out_dtype:
            y = y.to(x.dtype)
        return y.view(*x.shape[:-1], N), logq, logp

    @staticmethod
    @_guarded
    def backward(ctx, gy, glq, glp):
        lib = _lib.load()
        xg, W, w_mu, w_rho, b_mu, b_rho, wp_mu, wp_rho, bp_mu, bp_rho, z = ctx.saved_tensors
        spec, use_tc, M, x_dtype, x_shape, fused_act, use_x3, link_out = ctx.meta
        S = spec.S
        N, K = w_mu.shape
        dev = xg.device
        cdt = torch.bfloat16 if use_tc else torch.float32
        st = _stream(dev)
        if spec.kl_grad:
            glq, glp = _kl_upstream(glq, S, dev), _kl_upstream(glp, S, dev)
        else:
            glq = glp = None
        w_prior = PriorSpec(spec.w_prior.kind, spec.w_prior.pi, spec.w_prior.sigma1, spec.w_prior.sigma2, wp_mu, wp_rho)
        has_bias = b_mu is not None
        need_wmu, need_bmu = ctx.needs_input_grad[1], (has_bias and ctx.needs_input_grad[3])

        g_x = g_wmu = g_wrho = g_bmu = g_brho = None
        have_gy = gy is not None and M > 0
        db = None
        if have_gy:
            gyc = gy.reshape(S, M, N).to(cdt).contiguous()
            if z is not None and fused_act and link_out is not None and link_out.done:
                # the consumer's dgrad already multiplied by gelu'(z) (bf_linear_dgrad_gelu_bias): gy IS the gradient of
                # z, and its column sums -- this layer's bias gradient -- came out of the same epilogue
                link_out.done = False
                if link_out.bias_grad is not None and tuple(link_out.bias_grad.shape) == (S, N):
                    db = link_out.bias_grad
                link_out.bias_grad = None
            elif z is not None and fused_act:
                # gz = gy * gelu'(z) and the bias gradient's column sums in ONE pass over gy
                gz = torch.empty_like(gyc)
                db = torch.empty((S, N), dtype=torch.float32, device=dev)
                bws = _workspace("gelu_bwd", dev, lib.bf_gelu_bwd_bias_grad_workspace_bytes(S, M, N))
                rc = _timed("gelu_bwd_bias_grad", float(3 * S * M * N * 2), dev, lambda: lib.bf_gelu_bwd_bias_grad(
                    _ptr(gyc), _ptr(z), _ptr(gz), _ptr(db), S, M, N, _ptr(bws), st))
                _lib.check(rc, "bf_gelu_bwd_bias_grad")
                stats["launches"] += 1
                gyc = gz
            elif z is not None:
                gyc = torch.ops.aten.gelu_backward(gyc, z.reshape(S, M, N).to(cdt)).contiguous()
            gy_split = split_bf16x2(gyc) if use_x3 else None
            if ctx.needs_input_grad[0] and use_x3:
                dx = torch.empty((S, M, K), dtype=torch.float32, device=dev)
                rc = _timed("gemm_dgrad_x3", 2.0 * S * M * N * K, dev, lambda: lib.bf_linear_dgrad_x3(
                    _ptr(gy_split[0]), _ptr(gy_split[1]), _ptr(W[0]), _ptr(W[1]), _ptr(dx), S, M, N, K, st))
                _lib.check(rc, "bf_linear_dgrad_x3")
                stats["launches"] += 1
                g_x = dx.view(x_shape).to(x_dtype)
            elif (ctx.needs_input_grad[0] and use_tc and spec.gelu_in is not None and x_dtype == torch.bfloat16
                  and spec.gelu_in.z is not None and spec.gelu_in.z.numel() == S * M * K
                  and bool(lib.bf_linear_dgrad_gelu_supported(S, M, N, K))):
                # x = gelu(z) of a fused layer: dx o gelu'(z) straight from the dgrad epilogue, handed on as "dx"
                link = spec.gelu_in
                gz_in = torch.empty((S, M, K), dtype=torch.bfloat16, device=dev)
                db_in = torch.empty((S, K), dtype=torch.float32, device=dev)  # bias gradient of the layer that made z
                cws = _workspace("dgrad_gelu_colsum", dev, lib.bf_linear_dgrad_gelu_bias_workspace_bytes(S, M, K))
                rc = _timed("gemm_dgrad_tc", 2.0 * S * M * N * K, dev, lambda: lib.bf_linear_dgrad_gelu_bias(
                    _ptr(gyc), _ptr(W), _ptr(link.z), _ptr(gz_in), _ptr(db_in), _ptr(cws), S, M, N, K, st))
                _lib.check(rc, "bf_linear_dgrad_gelu_bias")
                stats["launches"] += 2  # the contraction and the fixed-order reduction of its per-block column sums
                g_x = gz_in.view(x_shape)
                link.done, link.buffer, link.z, link.bias_grad = True, g_x, None, db_in
            elif ctx.needs_input_grad[0]:
                dx_dtype = x_dtype if (use_tc and x_dtype in (torch.float32, torch.bfloat16)) else torch.float32
                sink = spec.sink
                buf = None if sink is None else sink.buffer
                if (buf is not None and use_tc and buf.dtype == dx_dtype == x_dtype and buf.is_contiguous()
                        and buf.numel() == S * M * K and buf.device == dev):
                    # the fused residual block already wrote its gradient of x there: add ours in place
                    rc = _timed("gemm_dgrad_tc", 2.0 * S * M * N * K, dev, lambda: lib.bf_linear_dgrad_accumulate(
                        _ptr(gyc), _ptr(W), _ptr(buf), S, M, N, K, _dt(cdt), _dt(dx_dtype), st))
                    _lib.check(rc, "bf_linear_dgrad_accumulate")
                    stats["launches"] += 1
                    stats["dgrad_accumulated"] += 1
                    sink.used = True
                    g_x = None
                else:
                    dx = torch.empty((S, M, K), dtype=dx_dtype, device=dev)
                    rc = _timed("gemm_dgrad_" + ("tc" if use_tc else "f32"), 2.0 * S * M * N * K, dev,
                                lambda: lib.bf_linear_dgrad(_ptr(gyc), _ptr(W), _ptr(dx), S, M, N, K, _dt(cdt),
                                                            _dt(dx_dtype), st))
                    _lib.check(rc, "bf_linear_dgrad")
                    stats["launches"] += 1
                    g_x = dx.view(x_shape).to(x_dtype)
            if has_bias and db is None and spec.bias_grad_box:
                # the consumer of y (fused dropout + residual + LayerNorm backward) already reduced gy over rows
                cand = spec.bias_grad_box.pop()
                spec.bias_grad_box.clear()
                if cand is not None and tuple(cand.shape) == (S, N) and cand.dtype == torch.float32:
                    db = cand
            if has_bias and db is None:
                db = torch.empty((S, N), dtype=torch.float32, device=dev)
                bws = _workspace("bias_grad", dev, lib.bf_bias_grad_workspace_bytes(S, M, N))
                rc = _timed("bias_grad", float(S * M * N * gyc.element_size()), dev,
                            lambda: lib.bf_bias_grad(_ptr(gyc), _dt(cdt), _ptr(db), S, M, N, _ptr(bws), st))
                _lib.check(rc, "bf_bias_grad")
                stats["launches"] += 1
            if use_tc:
                g_wrho = torch.empty_like(w_rho, dtype=torch.float32)
                g_wmu = torch.empty_like(w_rho, dtype=torch.float32) if need_wmu else None
                eps = _eps_arg(spec.w_stream, S, N * K, dev)
                ws = _workspace("wgrad_partials", dev,
                                lib.bf_linear_wgrad_fused_workspace_bytes(S, M, N, K, int(need_wmu)))
                pk = w_prior.kind
                rc = _timed("gemm_wgrad_fused_tc", 2.0 * S * M * N * K, dev, lambda: lib.bf_linear_wgrad_fused(
                    _ptr(gyc), _ptr(xg), S, M, N, K, BF_BF16, _ptr(w_mu), _ptr(w_rho), pk,
                    _ptr(wp_mu if pk == BF_PRIOR_GAUSSIAN else None),
                    _ptr(wp_rho if pk == BF_PRIOR_GAUSSIAN else None), w_prior.pi, w_prior.sigma1, w_prior.sigma2,
                    _ptr(glq), _ptr(glp), spec.w_stream.seed, spec.w_stream.step, spec.w_stream.tensor_id, _ptr(eps),
                    _ptr(g_wmu), _ptr(g_wrho), 0, _ptr(ws), st))
                _lib.check(rc, "bf_linear_wgrad_fused")
                stats["launches"] += 2  # contraction + fixed-order reduction
            else:
                dW = torch.empty((S, N, K), dtype=torch.float32, device=dev)
                if use_x3:
                    rc = _timed("gemm_wgrad_x3", 2.0 * S * M * N * K, dev, lambda: lib.bf_linear_wgrad_x3(
                        _ptr(gy_split[0]), _ptr(gy_split[1]), _ptr(xg[0]), _ptr(xg[1]), _ptr(dW), S, M, N, K, st))
                    _lib.check(rc, "bf_linear_wgrad_x3")
                else:
                    rc = _timed("gemm_wgrad_f32", 2.0 * S * M * N * K, dev,
                                lambda: lib.bf_linear_wgrad(_ptr(gyc), _ptr(xg), _ptr(dW), S, M, N, K, BF_F32, st))
                    _lib.check(rc, "bf_linear_wgrad")
                stats["launches"] += 1
                g_wmu, g_wrho = sample_kl_backward(dW.view(S, -1), w_mu.detach(), w_rho.detach(), w_prior,
                                                   spec.w_stream, S, glq, glp, need_wmu)
        else:
            g_wmu, g_wrho = sample_kl_backward(None, w_mu.detach(), w_rho.detach(), w_prior, spec.w_stream, S, glq,
                                               glp, need_wmu)
            db = None
        if has_bias:
            b_prior = PriorSpec(spec.b_prior.kind, spec.b_prior.pi, spec.b_prior.sigma1, spec.b_prior.sigma2,
                                bp_mu, bp_rho)
            g_bmu, g_brho = sample_kl_backward(db, b_mu.detach(), b_rho.detach(), b_prior, spec.b_stream, S, glq, glp,
                                               need_bmu)
        return g_x, g_wmu, g_wrho, g_bmu, g_brho, None, None, None, None, None


class LayerNormFn(torch.autograd.Function):
    """y = layer_norm(x) * gamma_s + beta_s for the S folded samples (rows
    [s*M, (s+1)*M) use sample s); gamma / beta are fp32 [S, H] (sampled) or [H]
    (shared affine).  Replaces F.layer_norm and its autograd."""

    @staticmethod
    @_guarded
    def forward(ctx, x, gamma, beta, S: int, eps: float):
        _require_cuda(x, "input")
        lib = _lib.load()
        dev = x.device
        H = x.shape[-1]
        rows = x.numel() // H
        if rows % S != 0:
            raise ValueError(f"{rows} rows are not a multiple of mc_samples={S}")
        M = rows // S
        xc = _aligned(x.detach())
        g = _aligned(gamma.detach().to(torch.float32))
        b = None if beta is None else _aligned(beta.detach().to(torch.float32))
        stride = H if g.dim() == 2 and g.shape[0] == S and S > 1 else 0
        if g.numel() != (S * H if stride else H):
            raise ValueError(f"affine of shape {tuple(gamma.shape)} does not match S={S}, H={H}")
        y = torch.empty_like(xc)
        mean = torch.empty(rows, dtype=torch.float32, device=dev)
        rstd = torch.empty(rows, dtype=torch.float32, device=dev)
        nbytes = float(rows * H * xc.element_size() * 2)
        rc = _timed("layernorm_fwd", nbytes, dev, lambda: lib.bf_layernorm_fwd(
            _ptr(xc), _dt(xc.dtype), _ptr(g), _ptr(b), stride, S, M, H, float(eps), _ptr(y), _ptr(mean), _ptr(rstd),
            _stream(dev)))
        _lib.check(rc, "bf_layernorm_fwd")
        stats["launches"] += 1
        ctx.save_for_backward(xc, g, mean, rstd)
        ctx.meta = (S, M, H, stride, gamma.shape, gamma.dtype, None if beta is None else beta.dtype)
        return y.view(x.shape)

    @staticmethod
    @_guarded
    def backward(ctx, gy):
        lib = _lib.load()
        xc, g, mean, rstd = ctx.saved_tensors
        S, M, H, stride, g_shape, g_dtype, b_dtype = ctx.meta
        dev = xc.device
        gyc = _aligned(gy.to(xc.dtype))
        dx = torch.empty_like(xc)
        Sa = S if stride else 1  # a shared affine is one "sample" spanning all rows
        dgamma = torch.empty((Sa, H), dtype=torch.float32, device=dev)
        dbeta = torch.empty((Sa, H), dtype=torch.float32, device=dev) if b_dtype is not None else None
        ws = _workspace("layernorm_bwd", dev, lib.bf_layernorm_bwd_workspace_bytes(Sa, M * S // Sa, H))
        nbytes = float(xc.numel() * xc.element_size() * 3)
        rc = _timed("layernorm_bwd", nbytes, dev, lambda: lib.bf_layernorm_bwd(
            _ptr(gyc), _ptr(xc), _dt(xc.dtype), _ptr(g), stride, _ptr(mean), _ptr(rstd), Sa, M * S // Sa, H, _ptr(dx),
            _ptr(dgamma), _ptr(dbeta), _ptr(ws), _stream(dev)))
        _lib.check(rc, "bf_layernorm_bwd")
        stats["launches"] += 1
        dg = dgamma.view(g_shape).to(g_dtype) if ctx.needs_input_grad[1] else None
        db = dbeta.view(g_shape).to(b_dtype) if (dbeta is not None and ctx.needs_input_grad[2]) else None
        return dx.view(gy.shape), dg, db, None, None


@dataclass
class DropoutSpec:
    """Identity of one dropout draw: keep mask = f(seed, site_id, step [+ device step], element)."""
    p: float = 0.0
    seed: int = 0
    site_id: int = 0
    step: int = 0


# ResidualLayerNormFn: hand the dropout keep bits from forward to backward (True) or regenerate them there (False: A/B, tests)
resln_keep_bits = {"on": True}


class ResidualLayerNormFn(torch.autograd.Function):
    """y = layer_norm(dropout_p(h) + r) * gamma_s + beta_s in one pass each way
    (`bf_resln_fwd` / `bf_resln_bwd`): the code around a Bayesian Linear in a
    transformer output block.  gamma / beta: fp32 [S, H] (sampled, row A10) or [H]
    (shared).  The dropout mask is a function of its Philox counter; its keep bits are handed to backward (128
    bytes per row) rather than regenerated there (`resln_keep_bits`).
    `bias_grad_box` (a list) receives sum_m dh[s][m][:] for the Linear that made h."""

    @staticmethod
    @_guarded
    def forward(ctx, h, r, gamma, beta, S: int, eps: float, drop: DropoutSpec, bias_grad_box, sink=None):
        _require_cuda(h, "input")
        lib = _lib.load()
        dev = h.device
        H = h.shape[-1]
        rows = h.numel() // H
        if r.shape != h.shape or r.dtype != h.dtype:
            raise ValueError(f"residual {tuple(r.shape)}/{r.dtype} does not match input {tuple(h.shape)}/{h.dtype}")
        if rows % S != 0:
            raise ValueError(f"{rows} rows are not a multiple of mc_samples={S}")
        M = rows // S
        hc, rc_ = _aligned(h.detach()), _aligned(r.detach())
        g = _aligned(gamma.detach().to(torch.float32))
        b = None if beta is None else _aligned(beta.detach().to(torch.float32))
        stride = H if g.dim() == 2 and g.shape[0] == S and S > 1 else 0
        if g.numel() != (S * H if stride else H):
            raise ValueError(f"affine of shape {tuple(gamma.shape)} does not match S={S}, H={H}")
        z, y = torch.empty_like(hc), torch.empty_like(hc)
        mean = torch.empty(rows, dtype=torch.float32, device=dev)
        rstd = torch.empty(rows, dtype=torch.float32, device=dev)
        nbytes = float(rows * H * hc.element_size() * 4)
        # the keep bits of the dropout mask go to backward (one word per lane and row: 128 B) instead of being regenerated there
        keep = torch.empty((rows, 32), dtype=torch.int32, device=dev) if (drop.p > 0 and resln_keep_bits["on"]) else None
        rc = _timed("resln_fwd", nbytes, dev, lambda: lib.bf_resln_fwd_keep(
            _ptr(hc), _ptr(rc_), _dt(hc.dtype), _ptr(g), _ptr(b), stride, S, M, H, float(eps), float(drop.p), drop.seed,
            drop.step & 0xFFFFFFFF, drop.site_id, _ptr(z), _ptr(y), _ptr(mean), _ptr(rstd), _ptr(keep), _stream(dev)))
        _lib.check(rc, "bf_resln_fwd_keep")
        stats["launches"] += 1
        ctx.save_for_backward(z, g, mean, rstd, keep)
        ctx.meta = (S, M, H, stride, gamma.shape, gamma.dtype, None if beta is None else beta.dtype, drop, bias_grad_box,
                    sink)
        return y.view(h.shape)

    @staticmethod
    @_guarded
    def backward(ctx, gy):
        lib = _lib.load()
        z, g, mean, rstd, keep = ctx.saved_tensors
        S, M, H, stride, g_shape, g_dtype, b_dtype, drop, box, sink = ctx.meta
        dev = z.device
        gyc = _aligned(gy.to(z.dtype))
        dz = torch.empty_like(z)
        dh = torch.empty_like(z) if drop.p > 0 else None
        Sa = S if stride else 1
        dgamma = torch.empty((Sa, H), dtype=torch.float32, device=dev)
        dbeta = torch.empty((Sa, H), dtype=torch.float32, device=dev) if b_dtype is not None else None
        dbias = torch.empty((S, H), dtype=torch.float32, device=dev) if box is not None else None
        ws = _workspace("resln_bwd", dev, lib.bf_resln_bwd_workspace_bytes(S, M, H))
        nbytes = float(z.numel() * z.element_size() * (4 if dh is not None else 3))
        rc = _timed("resln_bwd", nbytes, dev, lambda: lib.bf_resln_bwd_keep(
            _ptr(gyc), _ptr(z), _dt(z.dtype), _ptr(g), stride, _ptr(mean), _ptr(rstd), S, M, H, float(drop.p), drop.seed,
            drop.step & 0xFFFFFFFF, drop.site_id, _ptr(dz), _ptr(dh), _ptr(dgamma), _ptr(dbeta), _ptr(dbias), _ptr(ws),
            _ptr(keep), _stream(dev)))
        _lib.check(rc, "bf_resln_bwd_keep")
        stats["launches"] += 1
        if box is not None:
            box.clear()
            box.append(dbias)
        if dh is None and sink is not None:
            dh = dz.clone()  # dropout off: dz would alias the gradient of h, which must not see the accumulations
        dzv = dz.view(gy.shape)
        dhv = dzv if dh is None else dh.view(gy.shape)
        if sink is not None and ctx.needs_input_grad[1]:
            sink.buffer, sink.used = dzv, False  # Linear layers fed by the same input add their dgrad into it
        dg = dgamma.view(g_shape).to(g_dtype) if ctx.needs_input_grad[2] else None
        db = dbeta.view(g_shape).to(b_dtype) if (dbeta is not None and ctx.needs_input_grad[3]) else None
        return dhv, dzv, dg, db, None, None, None, None, None


class AttentionFn(torch.autograd.Function):
    """O = dropout_p(softmax(q k^T * scale)) v for short sequences (`bf_attention_fwd` / `_bwd`): the host model's
    attention between the Bayesian projections.  q, k, v: bf16 [B, H, T, 64] views with unit inner stride (HF passes
    transposed views of the [B, T, H*64] projection outputs: read in place); returns [B, T, H, 64].  The keep mask is
    regenerated in backward from (seed, site, step).

    `bias_boxes` (optional): three lists, the `bias_grad_box`es of the Bayesian q / k / v projections that produced the
    inputs, and `S` the number of folded samples: at T == 128 the backward kernel then also emits the column sums of dq,
    dk, dv per sample -- those layers' bias gradients -- and the layers skip their own pass over the gradients."""

    @staticmethod
    @_guarded
    def forward(ctx, q, k, v, scale: float, drop: DropoutSpec, bias_boxes=None, S: int = 1):
        _require_cuda(q, "query")
        lib = _lib.load()
        dev = q.device
        B, H, T, Dh = q.shape
        fix = lambda t: t if (t.stride(-1) == 1 and all(s % 8 == 0 for s in t.stride()[:3]) and t.data_ptr() % 16 == 0) \
            else t.contiguous()
        qc, kc, vc = fix(q.detach()), fix(k.detach()), fix(v.detach())
        strides = torch.tensor([t.stride(i) for t in (qc, kc, vc) for i in (0, 1, 2)], dtype=torch.int64)
        out = torch.empty((B, T, H, Dh), dtype=torch.bfloat16, device=dev)
        lse = torch.empty((B, H, T), dtype=torch.float32, device=dev)
        # T == 128 (tcgen05 kernels): the keep bits of the dropout mask go to backward instead of being regenerated
        keep = torch.empty((B, H, T, 4), dtype=torch.int32, device=dev) if (T == 128 and drop.p > 0.0) else None
        flops = 4.0 * B * H * T * T * Dh
        rc = _timed("attention_fwd", flops, dev, lambda: lib.bf_attention_fwd(
            _ptr(qc), _ptr(kc), _ptr(vc), strides.data_ptr(), B, H, T, float(scale), float(drop.p), drop.seed,
            drop.step & 0xFFFFFFFF, drop.site_id, _ptr(out), _ptr(lse), _ptr(keep), _stream(dev)))
        _lib.check(rc, "bf_attention_fwd")
        stats["launches"] += 1
        ctx.save_for_backward(qc, kc, vc, lse, keep)
        fused_bias = (bias_boxes is not None and T == 128 and (keep is not None or drop.p == 0.0) and S >= 1 and B % S == 0
                      and bool(lib.bf_get_option(_lib.BF_OPT_ATTN_TC)))
        ctx.meta = (float(scale), drop, strides, bias_boxes if fused_bias else None, int(S))
        return out

    @staticmethod
    @_guarded
    def backward(ctx, gout):
        lib = _lib.load()
        qc, kc, vc, lse, keep = ctx.saved_tensors
        scale, drop, strides, bias_boxes, S = ctx.meta
        dev = qc.device
        B, H, T, Dh = qc.shape
        g = gout.to(torch.bfloat16).contiguous()
        dq = torch.empty((B, T, H, Dh), dtype=torch.bfloat16, device=dev)
        dk, dv = torch.empty_like(dq), torch.empty_like(dq)
        flops = 10.0 * B * H * T * T * Dh
        dbias = ws = None
        if bias_boxes is not None:
            dbias = torch.empty((3, S, H * Dh), dtype=torch.float32, device=dev)
            ws = _workspace("attention_colsum", dev, lib.bf_attention_bias_workspace_bytes(B, H, S))
        rc = _timed("attention_bwd", flops, dev, lambda: lib.bf_attention_bwd_bias(
            _ptr(g), _ptr(qc), _ptr(kc), _ptr(vc), strides.data_ptr(), None, _ptr(lse), _ptr(keep), B, H, T, scale,
            float(drop.p), drop.seed, drop.step & 0xFFFFFFFF, drop.site_id, _ptr(dq), _ptr(dk), _ptr(dv), _ptr(dbias),
            _ptr(ws), S, _stream(dev)))
        _lib.check(rc, "bf_attention_bwd_bias")
        stats["launches"] += 1
        if bias_boxes is not None:
            stats["launches"] += 1  # the fixed-order reduction of the per-block column sums
            for i, box in enumerate(bias_boxes):  # q, k, v: picked up by the projections' backward (BayesLinear)
                if box is not None:
                    box.clear()
                    box.append(dbias[i])
        # [B, T, H, D] buffers seen as [B, H, T, D]: the layout the inputs came in (views of [B, T, H*D] rows), so the
        # transposes / reshapes of the surrounding model stay free
        return dq.transpose(1, 2), dk.transpose(1, 2), dv.transpose(1, 2), None, None, None, None


def attention_supported(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> bool:
    return (q.is_cuda and q.dtype == k.dtype == v.dtype == torch.bfloat16 and q.dim() == 4 and q.shape == k.shape == v.shape
            and bool(_lib.load().bf_attention_supported(int(q.shape[2]), int(q.shape[3]))))


def attention_dropout_mask(B: int, H: int, T: int, drop: DropoutSpec, device) -> torch.Tensor:
    """The keep mask `AttentionFn` applies, as uint8 [B, H, T, T] (1 = kept) -- tests."""
    lib = _lib.load()
    dev = torch.device(device)
    out = torch.empty((B, H, T, T), dtype=torch.uint8, device=dev)
    with on_device(dev):
        rc = lib.bf_attention_dropout_mask(_ptr(out), B, H, T, float(drop.p), drop.seed, drop.step & 0xFFFFFFFF,
                                           drop.site_id, _stream(dev))
    _lib.check(rc, "bf_attention_dropout_mask")
    return out


def dropout_mask(n: int, drop: DropoutSpec, device) -> torch.Tensor:
    """The keep mask (uint8, 1 = kept) `ResidualLayerNormFn` applies to a flat tensor of n elements -- tests."""
    lib = _lib.load()
    dev = torch.device(device)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    with on_device(dev):
        rc = lib.bf_dropout_mask(_ptr(out), n, float(drop.p), drop.seed, drop.step & 0xFFFFFFFF, drop.site_id,
                                 _stream(dev))
    _lib.check(rc, "bf_dropout_mask")
    return out


def resln_supported(h: torch.Tensor, r: torch.Tensor) -> bool:
    """Shapes / dtypes the fused dropout + residual + LayerNorm kernels take."""
    return (h.is_cuda and r.is_cuda and h.shape == r.shape and h.dtype == r.dtype
            and h.dtype in (torch.float32, torch.bfloat16) and bool(_lib.load().bf_resln_supported(int(h.shape[-1]))))


def layernorm_supported(x: torch.Tensor, normalized_shape) -> bool:
    """Shapes / dtypes the native LayerNorm kernels take (others use F.layer_norm)."""
    return (x.is_cuda and len(normalized_shape) == 1 and x.dtype in (torch.float32, torch.bfloat16)
            and bool(_lib.load().bf_layernorm_supported(int(normalized_shape[0]))))


def philox_normal(n: int, seed: int, step: int, tensor_id: int, sample_id: int, device) -> torch.Tensor:
    """eps stream of one (tensor, sample, step) -- for the statistical tests."""
    lib = _lib.load()
    dev = torch.device(device)
    out = torch.empty(n, dtype=torch.float32, device=dev)
    with on_device(dev):
        rc = lib.bf_philox_normal(_ptr(out), n, seed, step, tensor_id, sample_id, _stream(dev))
    _lib.check(rc, "bf_philox_normal")
    return out
