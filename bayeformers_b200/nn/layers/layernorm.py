"""Bayesian LayerNorm (absent from the reference snapshot).

`weight` and `bias` are `Gaussian((H,))`; forward samples both, accumulates
their log-probs into the layer scalars (weight then bias, as Linear does) and
applies layer normalisation with the sampled affine.  Under MOPED a
pretrained-style gamma = 1 gives sigma = delta and beta = 0 gives rho = 0
(sigma = ln 2) through the reference's -inf -> 0 rule.
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Size, Tensor

from ... import ops, runtime
from ..parameters.base import NoneParameter, Parameter
from ..parameters.gaussian import DEFAULT_SCALED_GAUSSIAN_MIXTURE, Gaussian, prior_spec_of
from ..parameters.initializations import DEFAULT_UNIFORM, Initialization
from .common import BayesianLayer, moped_


class LayerNorm(BayesianLayer):
    def __init__(self, normalized_shape: Union[int, Sequence[int]], eps: float = 1e-5,
                 elementwise_affine: bool = True, bias: bool = True,
                 initialization: Optional[Initialization] = DEFAULT_UNIFORM,
                 prior: Optional[Parameter] = DEFAULT_SCALED_GAUSSIAN_MIXTURE) -> None:
        super().__init__()
        if isinstance(normalized_shape, int):
            normalized_shape = (normalized_shape,)
        if not elementwise_affine:
            raise ValueError("a Bayesian LayerNorm needs elementwise_affine=True (there is nothing to sample otherwise)")
        self.normalized_shape = tuple(normalized_shape)
        self.eps = eps
        self.initialization = initialization
        self.weight = Gaussian(Size(self.normalized_shape), self.initialization)
        self.weight_prior = prior
        if bias:
            self.bias = Gaussian(Size(self.normalized_shape), self.initialization)
            self.bias_prior = prior
        else:
            self.bias = NoneParameter()
            self.bias_prior = NoneParameter()
        self._init_scalars()

    def sample_affine(self):
        """Draw gamma_s / beta_s for the current mc_samples and publish the layer's log-probs
        (weight then bias, as Linear does).  Returns (w [S, H] fp32, b [S, H] fp32 or None, S).  One autograd node
        (`ops.SamplePair`); the draw comes from the multi-tensor sampler when `enable_presample` is on."""
        S = runtime.get_mc_samples()
        kl_grad = self._kl_grad()
        has_bias = isinstance(self.bias, Gaussian)
        wp = prior_spec_of(self.weight_prior)
        bp = prior_spec_of(self.bias_prior) if has_bias else ops.PriorSpec()
        pre, self._presampled = self._presampled, None  # a draw serves exactly one forward
        if pre is not None and pre[0] != S:
            pre = None
        if pre is not None:
            spec = ops.PairSpec(S=S, kl_grad=kl_grad, w_prior=wp, b_prior=bp, w_stream=pre[5], b_stream=pre[6],
                                presampled=pre[1:5])
        else:
            spec = ops.PairSpec(S=S, kl_grad=kl_grad, w_prior=wp, b_prior=bp, w_stream=self.weight.next_stream(S),
                                b_stream=self.bias.next_stream(S) if has_bias else ops.StreamSpec())
        self._last_streams = (spec.w_stream, spec.b_stream)
        w, b, logq, logp = ops.SamplePair.apply(
            self.weight.mu, self.weight.rho, self.bias.mu if has_bias else None, self.bias.rho if has_bias else None,
            wp.mu, wp.rho, bp.mu, bp.rho, spec)
        self._publish(logq, logp, S, kl_grad, means=pre[7] if pre is not None else None)
        return w, (b if has_bias else None), S

    def forward(self, input: Tensor) -> Tensor:
        w, b, S = self.sample_affine()
        if len(self.normalized_shape) == 1 and ops.layernorm_supported(input, self.normalized_shape):
            if input.shape[0] % S != 0:
                raise ValueError(f"leading dimension {input.shape[0]} is not a multiple of mc_samples={S}")
            # native S-sample kernel: statistics, per-sample affine and their gradients in one pass each way
            return ops.LayerNormFn.apply(input, w, b, S, self.eps)
        if S == 1:
            return F.layer_norm(input, self.normalized_shape, w[0].to(input.dtype),
                                None if b is None else b[0].to(input.dtype), self.eps)
        if input.shape[0] % S != 0:
            raise ValueError(f"leading dimension {input.shape[0]} is not a multiple of mc_samples={S}")
        xn = F.layer_norm(input, self.normalized_shape, None, None, self.eps)
        xs = xn.view(S, input.shape[0] // S, *input.shape[1:])
        shape_w = [S, 1] + [1] * (input.dim() - 1 - len(self.normalized_shape)) + list(self.normalized_shape)
        y = xs * w.to(input.dtype).view(shape_w)
        if b is not None:
            y = y + b.to(input.dtype).view(shape_w)
        return y.view(input.shape)

    @classmethod
    def from_frequentist(cls, ln: nn.Module, initialization: Optional[Initialization] = DEFAULT_UNIFORM,
                         prior: Optional[Parameter] = DEFAULT_SCALED_GAUSSIAN_MIXTURE, delta: float = None,
                         freeze: bool = False) -> "LayerNorm":
        if ln.weight is None:
            raise ValueError("cannot convert a LayerNorm without affine parameters")
        has_bias = ln.bias is not None
        baye = cls(ln.normalized_shape, ln.eps, True, has_bias, prior=prior)
        if delta is not None:
            baye.weight_prior = moped_(baye.weight, ln.weight, delta, freeze)
            if has_bias:
                baye.bias_prior = moped_(baye.bias, ln.bias, delta, freeze)
        return baye


class HostLayerNorm(nn.LayerNorm):
    """Frequentist `nn.LayerNorm` of the host model routed through the same
    native kernels (shared affine, fp32 master gamma/beta, activations in the
    input's dtype).  Installed by `bayeformers_b200.accelerate_host_`; numerics
    are those of F.layer_norm with fp32 statistics.  Falls back to the stock
    implementation for shapes the kernels do not take."""

    def forward(self, input: Tensor) -> Tensor:
        if (self.weight is not None and len(self.normalized_shape) == 1
                and ops.layernorm_supported(input, self.normalized_shape)):
            return ops.LayerNormFn.apply(input, self.weight, self.bias, 1, self.eps)
        return super().forward(input)
