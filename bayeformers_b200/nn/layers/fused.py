"""Fused transformer "output" block around a Bayesian Linear.

HuggingFace-style blocks of the form (transformers' BertSelfOutput / BertOutput and
their clones)

    def forward(self, hidden_states, input_tensor):
        hidden_states = self.dense(hidden_states)
        hidden_states = self.dropout(hidden_states)
        return self.LayerNorm(hidden_states + input_tensor)

are re-routed, in place, by `bayeformers_b200.accelerate_host_(..., fuse_residual=True)`:
`dense` stays the Bayesian Linear it was (`bnn.Linear`, reference
bayeformers/nn/layers/linear.py:83-104); dropout + residual add + LayerNorm become one
kernel pass each way (`ops.ResidualLayerNormFn`), whose backward also hands `dense` the
row sums of its output gradient, i.e. its bias gradient.  The LayerNorm may be
frequentist (`nn.LayerNorm` / `HostLayerNorm`: shared affine) or Bayesian
(`bnn.LayerNorm`, SURVEY.md row A10: per-sample affine, log-probs published as usual).
Parameters, state_dict names and results (to rounding; the dropout mask comes from the
Philox stream instead of torch's generator) are unchanged.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from ... import ops, runtime
from ..parameters.gaussian import Gaussian
from .layernorm import LayerNorm as BayesLayerNorm
from .linear import Linear

_FUSED_CLASSES: Dict[type, type] = {}


class FusedOutputMixin:
    """forward(hidden_states, input_tensor) of a dense -> dropout -> LayerNorm(+ residual) block."""

    def forward(self, hidden_states: torch.Tensor, input_tensor: torch.Tensor) -> torch.Tensor:
        dense, ln, drop = self.dense, self.LayerNorm, self.dropout
        fusable = (isinstance(dense, Linear) and input_tensor.is_cuda and len(ln.normalized_shape) == 1
                   and ln.normalized_shape[0] == dense.out_features and getattr(ln, "weight", None) is not None)
        if not fusable:
            return super().forward(hidden_states, input_tensor)
        box = [] if (isinstance(dense.bias, Gaussian) and torch.is_grad_enabled()) else None
        dense._bias_grad_box = box
        h = dense(hidden_states)
        if not ops.resln_supported(h, input_tensor):
            return ln(drop(h) + input_tensor)
        if isinstance(ln, BayesLayerNorm):
            gamma, beta, S = ln.sample_affine()
        else:
            gamma, beta, S = ln.weight, ln.bias, runtime.get_mc_samples()
        if (h.numel() // h.shape[-1]) % S != 0:
            raise ValueError(f"{h.numel() // h.shape[-1]} rows are not a multiple of mc_samples={S}")
        p = float(drop.p) if (drop.training and drop.p > 0) else 0.0
        self._bf_calls = getattr(self, "_bf_calls", 0) + 1
        spec = ops.DropoutSpec(p=p, seed=runtime.dropout_seed(), site_id=self._bf_site, step=self._bf_calls & 0xFFFFFFFF)
        self._last_dropout = spec  # identity of this forward's mask (tests, debugging)
        sink = runtime.sink_for(input_tensor, create=False) if runtime.grad_sinks_enabled() else None
        return ops.ResidualLayerNormFn.apply(h, input_tensor, gamma, beta, S, ln.eps, spec, box, sink)


def is_output_block(mod: nn.Module) -> bool:
    children = dict(mod.named_children())
    return (set(children) == {"dense", "LayerNorm", "dropout"} and isinstance(children["dense"], Linear)
            and isinstance(children["dropout"], nn.Dropout)
            and isinstance(children["LayerNorm"], (nn.LayerNorm, BayesLayerNorm))
            and not isinstance(mod, FusedOutputMixin))


def fuse_output_block_(mod: nn.Module) -> nn.Module:
    cls = mod.__class__
    fused = _FUSED_CLASSES.get(cls)
    if fused is None:
        fused = type("Fused" + cls.__name__, (FusedOutputMixin, cls), {"__module__": __name__})
        _FUSED_CLASSES[cls] = fused
        globals()[fused.__name__] = fused  # importable by name (deepcopy / pickle of the patched model)
    mod.__class__ = fused
    mod._bf_site = runtime.next_tensor_id()
    mod._bf_calls = 0
    return mod
