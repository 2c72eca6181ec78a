"""bayeformers_b200 -- B200-native drop-in for the variational-layer hot path of
yliess86/BayeFormers.

    from bayeformers_b200 import to_bayesian
    import bayeformers_b200.nn as bnn

mirrors `from bayeformers import to_bayesian` / `import bayeformers.nn as bnn`.
"""
from __future__ import annotations

from copy import deepcopy
from typing import Dict, Optional

import torch
import torch.nn as tnn

from . import nn, optim, runtime  # noqa: F401  (bayeformers_b200.nn is part of the public surface)
from .nn import TORCH2BAYE, TORCH2BAYE_ALL
from .nn.model import Model
from .nn.parameters.base import Parameter
from .nn.parameters.gaussian import DEFAULT_SCALED_GAUSSIAN_MIXTURE
from .nn.parameters.initializations import DEFAULT_UNIFORM, Initialization
from .harness import predictive_stats, sample_bayesian
from .presample import enable_presample
from .runtime import (advance_step, disable_device_step, enable_device_step, hf_capture_compat, load_rng_state,
                      manual_seed, mc_samples, rng_state, set_gemm_dtype, set_kl_grad)

__all__ = ["to_bayesian", "cast_frequentist_", "accelerate_host_", "enable_presample", "sample_bayesian",
           "predictive_stats", "nn", "Model", "manual_seed", "mc_samples", "set_gemm_dtype",
           "set_kl_grad", "enable_device_step", "disable_device_step", "advance_step", "hf_capture_compat", "rng_state",
           "load_rng_state"]


def cast_frequentist_(model: tnn.Module, dtype: torch.dtype) -> tnn.Module:
    """Cast, in place, the floating-point parameters and buffers of every module
    that is NOT part of a Bayesian layer (embeddings, LayerNorm, ... of the host
    model) to `dtype`, so activations flow in bf16 when `gemm_dtype="bf16"`.
    The variational masters (mu, rho, priors) always stay fp32 (quirk Q4)."""
    from .nn.layers.common import BayesianLayer

    from .nn.layers.layernorm import HostLayerNorm

    def walk(mod: tnn.Module) -> None:
        if isinstance(mod, (BayesianLayer, HostLayerNorm)):  # fp32 masters stay fp32
            return
        for p in mod._parameters.values():
            if p is not None and p.is_floating_point():
                p.data = p.data.to(dtype)
        for name, b in mod._buffers.items():
            if b is not None and b.is_floating_point():
                mod._buffers[name] = b.to(dtype)
        for child in mod.children():
            walk(child)

    walk(model)
    return model


def _is_exact_gelu(fn) -> bool:
    if isinstance(fn, tnn.GELU):
        return getattr(fn, "approximate", "none") == "none"
    name = type(fn).__name__
    if name == "GELUActivation":  # transformers.activations: exact erf GELU unless use_gelu_python
        return getattr(fn, "act", None) is tnn.functional.gelu
    return fn is tnn.functional.gelu


def accelerate_host_(model: tnn.Module, layernorm: bool = True, fuse_gelu: bool = True,
                     fuse_residual: bool = False, grad_sinks: bool = False, attention: bool = False,
                     attention_bias_grads: bool = False) -> tnn.Module:
    """Opt-in plumbing for the frequentist code AROUND the Bayesian layers of a host model.  In place; parameters,
    state_dict names and numerics (to rounding) are unchanged.

    layernorm  route the remaining frequentist `nn.LayerNorm`s (exact class, 1-D normalized_shape) through the
               native S-sample LayerNorm kernels with a shared affine (fp32 master gamma / beta).
    fuse_gelu  HuggingFace feed-forward blocks of the form `x = self.dense(x); x = self.intermediate_act_fn(x)`
               (BertIntermediate and its clones) whose `dense` is a Bayesian Linear and whose activation is the
               exact GELU: move the activation INTO the layer (`dense.activation = "gelu"`, fused tensor-core
               epilogue) and replace the module's activation by the identity.
    fuse_residual  HuggingFace output blocks `LayerNorm(dropout(dense(x)) + input_tensor)` (BertSelfOutput,
               BertOutput and their clones) whose `dense` is a Bayesian Linear: dropout + residual add + LayerNorm
               run as one kernel pass each way, the dropout mask comes from the Philox counter stream (never
               stored), and the pass also yields `dense`'s bias gradient (nn/layers/fused.py).
    grad_sinks (needs fuse_residual; process-wide switch) an input that feeds Bayesian Linear layers and the residual
               of a fused output block gets its gradient accumulated in place: the block's backward writes its part
               first, the Linear dgrad kernels add theirs into that buffer by TMA reduce-add, and autograd's separate
               add passes disappear.  Valid when such an input has no further consumers (HF BERT); a tensor hook
               raises otherwise (runtime.GradSink).
    attention  transformers models: route self-attention over short sequences (bf16, no mask, T <= 128, head width 64)
               through the native whole-sequence kernels (nn/layers/attention.py); other cases keep torch's SDPA.  The
               attention dropout mask then comes from the Philox counter stream.
    attention_bias_grads  with `attention`: the T = 128 backward kernel also emits the bias gradients of the Bayesian query /
               key / value projections (column sums of dq, dk, dv) and those layers skip their own pass.  Off by default:
               measured at the bench shape the two store warps' extra instructions cost the backward kernel what the
               three skipped passes save (attention_bwd 10.0 -> 13.2 ms against 3.0 ms of bias_grad per step)."""
    from .nn.layers.fused import fuse_output_block_, is_output_block
    from .nn.layers.layernorm import HostLayerNorm
    from .nn.layers.linear import Linear

    if grad_sinks and not fuse_residual:
        raise ValueError("grad_sinks=True needs fuse_residual=True")
    if attention:
        from .nn.layers.attention import use_native_attention_
        use_native_attention_(model, bias_grads=attention_bias_grads)
    if grad_sinks:
        runtime.enable_grad_sinks(True)  # process-wide; runtime.enable_grad_sinks(False) switches it off again
    if fuse_residual:
        for mod in list(model.modules()):
            if is_output_block(mod):
                fuse_output_block_(mod)
    for mod in model.modules():
        if layernorm and mod.__class__ is tnn.LayerNorm and mod.elementwise_affine and len(mod.normalized_shape) == 1:
            mod.__class__ = HostLayerNorm
        if fuse_gelu and isinstance(getattr(mod, "dense", None), Linear) and hasattr(mod, "intermediate_act_fn"):
            children = dict(mod.named_children())
            only_dense_and_act = set(children) <= {"dense", "intermediate_act_fn"}
            if only_dense_and_act and mod.dense.activation is None and _is_exact_gelu(mod.intermediate_act_fn):
                mod.dense.activation = "gelu"
                mod.intermediate_act_fn = tnn.Identity()
    return model


def to_bayesian(model: tnn.Module, initialization: Optional[Initialization] = DEFAULT_UNIFORM,
                prior: Optional[Parameter] = DEFAULT_SCALED_GAUSSIAN_MIXTURE, delta: float = None,
                freeze: bool = False, *, layers: Optional[Dict[type, type]] = None, gemm_dtype=None,
                kl_grad: Optional[bool] = None) -> Model:
    """Deep-copy `model` and swap every convertible child for its Bayesian
    equivalent; returns the copy wrapped in `bnn.Model`.

    Same positional contract as the reference
    (/root/reference/bayeformers/__init__.py:19-63): exact-class matching (so
    subclasses and the root module are left alone, quirk Q7), `delta` switches
    MOPED initialisation on, `freeze` stops gradients to mu.

    Keyword-only extensions (defaults = reference behaviour):
      layers      registry to use; default `TORCH2BAYE` (nn.Linear only, like the
                  reference).  Pass `bnn.TORCH2BAYE_ALL` to convert Embedding and
                  LayerNorm as well.
      gemm_dtype  "fp32" (reference precision, FFMA kernels), "bf16" (tcgen05 tensor cores, 1e-2) or "fp32x3"
                  (reference precision ON the tensor cores: 3-pass bf16 split, 1e-5)
      kl_grad     True lets the KL term (log q - log p) back-propagate; the
                  reference silently drops it (linear.py:99-102)
    """
    registry = TORCH2BAYE if layers is None else layers
    dt = None if gemm_dtype is None else runtime._as_dtype(gemm_dtype)

    def replace(parent: tnn.Module) -> None:
        for name, child in parent.named_children():
            if child.__class__ in registry:
                baye = registry[child.__class__].from_frequentist(child, initialization, prior, delta, freeze)
                baye.gemm_dtype, baye.kl_grad = dt, kl_grad
                setattr(parent, name, baye)
            else:
                replace(child)

    new_model = deepcopy(model)
    replace(new_model)
    return Model(model=new_model)
