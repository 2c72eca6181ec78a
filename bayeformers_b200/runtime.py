"""Process-wide runtime options of the variational layers.

Everything here is an *additive extension* over the reference API; the
defaults reproduce the reference's behaviour:

  mc_samples = 1      one weight sample per forward (bayeformers/nn/layers/linear.py:97)
  kl_grad    = False  log-probs detached like the reference's `.data =` (linear.py:99-102)
  gemm_dtype = fp32   reference precision (bayeformers/nn/parameters/base.py:32)

The eps stream is counter based (Philox4x32-10): eps is a pure function of
(seed, tensor_id, step, sample_id, element), so backward regenerates it and
data-parallel replicas draw identical weights with no communication.
"""
from __future__ import annotations

import contextlib
import itertools
import threading
from typing import Optional

import torch

_state = threading.local()
_tensor_ids = itertools.count(1)
_global = {"seed": None, "kl_grad": False, "gemm_dtype": torch.float32, "dropout_salt": 0}


def next_tensor_id() -> int:
    """Stream id of a new variational tensor (construction order => identical
    on every rank that builds the same model)."""
    return next(_tensor_ids)


def manual_seed(seed: int) -> None:
    """Seed of the eps stream.  Call with the same value on every rank."""
    _global["seed"] = int(seed) & 0xFFFFFFFFFFFFFFFF


def seed() -> int:
    """Current seed; if never set it is taken once from torch's default
    generator, so `torch.manual_seed` also fixes the eps stream."""
    if _global["seed"] is None:
        _global["seed"] = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
    return _global["seed"]


def set_dropout_salt(salt: int) -> None:
    """Per-rank salt of the DROPOUT streams (fused output blocks, native attention).  Under batch sharding every rank
    must draw the same eps (same seed) but should NOT apply the same dropout mask to its own rows: `parallel.
    broadcast_seed` sets the salt to the rank, which changes the masks and leaves the weight samples alone."""
    _global["dropout_salt"] = int(salt)


def dropout_seed() -> int:
    """Seed of the counter-based dropout masks: the eps seed mixed with the per-rank salt."""
    return (seed() + 0x9E3779B97F4A7C15 * _global["dropout_salt"]) & 0xFFFFFFFFFFFFFFFF


def get_mc_samples() -> int:
    return getattr(_state, "mc_samples", 1)


@contextlib.contextmanager
def mc_samples(S: int):
    """Fold S Monte-Carlo samples into the leading (batch) dimension: inside
    the context a Bayesian layer treats its input as [S*B, ...], uses weight
    sample s for rows [s*B, (s+1)*B) and reports [S]-shaped log-probs.
    Bit-identical to the reference's sequential S-loop given the same eps
    (SURVEY.md section 0, item 3)."""
    S = int(S)
    if S < 1:
        raise ValueError("mc_samples must be >= 1")
    prev = get_mc_samples()
    _state.mc_samples = S
    try:
        yield
    finally:
        _state.mc_samples = prev


def set_folded_rows(rows: Optional[int]) -> None:
    """Leading dimension (S*B) of the folded batch of the forward in flight; set by `bnn.Model.forward`
    so that layers fed a broadcast input of leading dimension 1 (HuggingFace passes `position_ids` of
    shape [1, T] to its position embedding) can expand it to the folded batch."""
    _state.folded_rows = rows


def get_folded_rows() -> Optional[int]:
    return getattr(_state, "folded_rows", None)


# ---- gradient sinks (opt-in, accelerate_host_(grad_sinks=True)) ---------------------------------------------
class GradSink:
    """One per tensor x that feeds Bayesian Linear layers AND the residual input of a fused output block in the
    same forward.  Backward: the fused block's kernel writes its gradient of x (dz) first and parks it here; the
    Linear layers then ADD their dgrad results into that buffer in place (TMA reduce-add) and hand autograd no
    gradient of their own, so the engine never runs its separate `grad + grad` passes for x."""
    __slots__ = ("tensor", "buffer", "used")

    def __init__(self, tensor: torch.Tensor) -> None:
        self.tensor, self.buffer, self.used = tensor, None, False

    def check(self, grad: torch.Tensor) -> None:
        """Tensor hook on x: the gradient autograd finally delivers for x must be the buffer the Linear layers
        accumulated into -- anything else means x had a consumer this scheme does not know (its contribution made
        the engine build a new sum, and later in-place accumulations were lost)."""
        buf, used = self.buffer, self.used
        self.buffer, self.used, self.tensor = None, False, None
        if used and (grad is None or buf is None or grad.data_ptr() != buf.data_ptr()):
            raise RuntimeError("bayeformers_b200: gradient sinks are not valid for this model (an input shared by "
                               "Bayesian Linear layers and a fused residual has further consumers); call "
                               "accelerate_host_(..., grad_sinks=False)")


class GeluLink:
    """Hand-off between a Bayesian Linear with a fused GELU (producer of a = gelu(z)) and the Bayesian Linear that
    consumes a.  Attached to the tensor a by the producer's forward.  The consumer's backward multiplies its input
    gradient by gelu'(z) inside the dgrad kernel (`bf_linear_dgrad_gelu`) and returns that product AS the gradient of
    a; `done` tells the producer's backward that what arrives is already the gradient of z, so it skips its own GELU'
    pass.  Valid when a has no other consumer (HF BertIntermediate -> BertOutput): a tensor hook on a checks that the
    gradient autograd finally delivers is exactly the buffer the fused kernel wrote, and raises otherwise."""
    __slots__ = ("z", "done", "buffer", "bias_grad")

    def __init__(self, z: torch.Tensor) -> None:
        self.z, self.done, self.buffer = z, False, None
        self.bias_grad = None  # [S, N] fp32 column sums of the gradient of z, when the consumer's dgrad kernel emitted them

    def check(self, grad: torch.Tensor) -> None:
        if self.done and (grad is None or self.buffer is None or grad.data_ptr() != self.buffer.data_ptr()):
            raise RuntimeError("bayeformers_b200: the output of a GELU-fused Bayesian Linear feeds more than the next "
                               "Bayesian Linear, so the fused GELU' dgrad is not valid for this model; call "
                               "bayeformers_b200.runtime.enable_gelu_links(False)")
        self.buffer = None


_gelu_links = {"on": True}


def enable_gelu_links(flag: bool = True) -> None:
    """Process-wide switch of the fused GELU' dgrad (on by default; only ever active behind a fused-GELU layer)."""
    _gelu_links["on"] = bool(flag)


def gelu_links_enabled() -> bool:
    return _gelu_links["on"]


_sinks_enabled = {"on": False}


def enable_grad_sinks(flag: bool = True) -> None:
    _sinks_enabled["on"] = bool(flag)


def grad_sinks_enabled() -> bool:
    return _sinks_enabled["on"]


def reset_sinks() -> None:
    _state.sinks = {}


def sink_for(x: torch.Tensor, create: bool) -> Optional[GradSink]:
    """The GradSink of tensor `x` in the forward in flight (keyed by identity; the sink keeps x alive, so ids are
    not reused within one forward)."""
    table = getattr(_state, "sinks", None)
    if table is None:
        table = _state.sinks = {}
    sink = table.get(id(x))
    if sink is None and create:
        sink = table[id(x)] = GradSink(x)
        x.register_hook(sink.check)
    return sink


def set_kl_grad(flag: bool) -> None:
    _global["kl_grad"] = bool(flag)


def get_kl_grad() -> bool:
    return _global["kl_grad"]


FP32X3 = "fp32x3"  # reference precision on the tensor cores: 3-pass bf16 (hi, lo) split contractions


def _as_dtype(d):
    """torch.float32 (FFMA parity kernels), torch.bfloat16 (tcgen05, 1e-2) or FP32X3 (tcgen05 split precision, 1e-5)."""
    if isinstance(d, torch.dtype):
        out = d
    else:
        out = {"fp32": torch.float32, "float32": torch.float32, "bf16": torch.bfloat16,
               "bfloat16": torch.bfloat16, "fp32x3": FP32X3}[str(d)]
    if out not in (torch.float32, torch.bfloat16, FP32X3):
        raise ValueError("gemm_dtype must be fp32, bf16 or fp32x3")
    return out


def activation_dtype(d) -> torch.dtype:
    """dtype activations / sampled rows travel in under GEMM mode `d`."""
    return torch.bfloat16 if d == torch.bfloat16 else torch.float32


def set_gemm_dtype(d) -> None:
    _global["gemm_dtype"] = _as_dtype(d)


def get_gemm_dtype() -> torch.dtype:
    return _global["gemm_dtype"]


# ---- device-resident step counter (CUDA-graph capture of a whole training step) ----------
_device_step = {}  # device index -> int32[1] tensor registered with the library for that device


def enable_device_step(device) -> torch.Tensor:
    """Allocate the step counter of `device` and register it with the kernels.  From
    then on eps is a function of (seed, tensor_id, host_step + device_step,
    sample): call `advance_step()` once per training step -- it is a plain
    device add, so it can live inside a captured CUDA graph."""
    from . import _lib, ops

    dev = torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    t = torch.zeros(1, dtype=torch.int32, device=dev)
    with ops.on_device(dev):
        _lib.check(_lib.load().bf_set_step_counter(t.data_ptr()), "bf_set_step_counter")
    _device_step[dev.index] = t
    return t


def disable_device_step(device=None) -> None:
    from . import _lib, ops

    for idx in list(_device_step) if device is None else [torch.device(device).index]:
        if idx in _device_step:
            with ops.on_device(torch.device("cuda", idx)):
                _lib.check(_lib.load().bf_set_step_counter(None), "bf_set_step_counter")
            del _device_step[idx]


def advance_step(n: int = 1, device=None) -> None:
    if not _device_step:
        raise RuntimeError("enable_device_step(device) first")
    if device is None:
        for t in _device_step.values():
            t.add_(n)
    else:
        _device_step[torch.device(device).index].add_(n)


# ---- reproducible resume (checkpoint row) -----------------------------------------------
def rng_state(model: torch.nn.Module) -> dict:
    """Everything that determines the NEXT eps draw and dropout mask of `model`: the Philox seed, the stream id and
    the per-tensor step of every variational tensor, the multi-tensor sampler's run counter, the call counters of the
    fused dropout sites and the device step counters.  Plain Python / CPU values: store it NEXT TO the state_dict

        torch.save({"model": bm.state_dict(), "rng": bf.rng_state(bm), "optimizer": opt.state_dict()}, path)

    (the reference's checkpoints hold `"model": b_model.state_dict()` plus scalars, examples/bert_glue.py:303-309).
    It is deliberately not part of `state_dict()`: an extra key there would break strict loading of our checkpoints
    into the reference and of reference checkpoints into this package.  Reads the device counters (one sync)."""
    from .nn.parameters.gaussian import Gaussian

    out = {"version": 1, "seed": seed(), "dropout_salt": _global["dropout_salt"], "gaussians": {}, "dropout_sites": {},
           "presample_runs": None,
           "device_steps": {int(i): int(t.item()) for i, t in _device_step.items()}}
    for name, mod in model.named_modules():
        if isinstance(mod, Gaussian):
            out["gaussians"][name] = {"tensor_id": int(mod.tensor_id), "step": int(mod.step)}
        if hasattr(mod, "_bf_site"):  # fused output blocks and native-attention modules
            out["dropout_sites"][name] = {"site": int(mod._bf_site), "calls": int(getattr(mod, "_bf_calls", 0))}
    ps = getattr(model, "_presampler", None)
    if ps is not None:
        out["presample_runs"] = int(ps._runs)
    return out


def load_rng_state(model: torch.nn.Module, state: dict) -> None:
    """Inverse of `rng_state`: after this the model's next forward draws exactly what the checkpointed process
    would have drawn next (same seed, stream ids, steps and dropout counters)."""
    from .nn.parameters.gaussian import Gaussian

    if state.get("version") != 1:
        raise ValueError(f"unknown rng_state version {state.get('version')!r}")
    manual_seed(int(state["seed"]))
    set_dropout_salt(int(state.get("dropout_salt", 0)))
    mods = dict(model.named_modules())
    for name, g in state["gaussians"].items():
        mod = mods.get(name)
        if not isinstance(mod, Gaussian):
            raise KeyError(f"rng_state names a variational tensor {name!r} this model does not have")
        mod.tensor_id, mod.step = int(g["tensor_id"]), int(g["step"])
    for name, d in state["dropout_sites"].items():
        mod = mods.get(name)
        if mod is None or not hasattr(mod, "_bf_site"):
            raise KeyError(f"rng_state names a fused dropout site {name!r} this model does not have "
                           "(call accelerate_host_ with the same options before loading)")
        mod._bf_site, mod._bf_calls = int(d["site"]), int(d["calls"])
    ps = getattr(model, "_presampler", None)
    if state.get("presample_runs") is not None:
        if ps is None:
            raise KeyError("rng_state was taken with multi-tensor sampling on: call enable_presample(model) first")
        ps._runs = int(state["presample_runs"])
        ps._sig = None  # stream ids may have changed: rebuild the descriptor table
    for idx, value in state.get("device_steps", {}).items():
        t = _device_step.get(int(idx))
        if t is None:
            raise KeyError(f"rng_state holds a device step for cuda:{idx}: call enable_device_step first")
        t.fill_(int(value))


@contextlib.contextmanager
def hf_capture_compat():
    """Use around `torch.cuda.graph(...)` when the host model is a HuggingFace transformer.

    transformers treats "the CUDA stream is capturing" like tracing and then always
    materialises a [B,1,T,T] attention mask (transformers/masking_utils.py,
    `_ignore_bidirectional_mask_sdpa`), which pushes SDPA from the fused cuDNN kernel
    onto the unfused math path (and into cuBLAS calls that cannot initialise while
    capturing).  With `attention_mask=None` there is no data-dependent branch to
    protect, so inside this context the eager behaviour (no mask) is kept.  No-op
    when transformers is not importable."""
    try:
        import transformers.masking_utils as hf_masking
    except Exception:  # pragma: no cover
        yield
        return
    saved = getattr(hf_masking, "is_tracing", None)
    if saved is None:
        yield
        return
    hf_masking.is_tracing = lambda *a, **k: False
    try:
        yield
    finally:
        hf_masking.is_tracing = saved
