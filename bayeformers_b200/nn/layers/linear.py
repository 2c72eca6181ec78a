"""Bayesian Linear on B200.

API parity with /root/reference/bayeformers/nn/layers/linear.py:24-164: same
constructor, same attributes (`weight`, `bias`, `weight_prior`, `bias_prior`,
`log_prior`, `log_variational_posterior`), same `from_frequentist` classmethod
and the same state_dict names.  `forward` is one autograd node
(`ops.BayesLinear`) instead of ~100 torch kernels: fused Philox sample + log q
+ log p, then the S-sample contraction on tensor cores, with eps recomputed in
backward and dmu/drho produced in the weight-gradient epilogue.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch import Size, Tensor

from ... import ops, runtime
from ..parameters.base import NoneParameter, Parameter
from ..parameters.gaussian import DEFAULT_SCALED_GAUSSIAN_MIXTURE, Gaussian, prior_spec_of
from ..parameters.initializations import DEFAULT_UNIFORM, Initialization
from .common import BayesianLayer, moped_


class Linear(BayesianLayer):
    def __init__(self, in_features: int, out_features: int, bias: Optional[bool] = True,
                 initialization: Optional[Initialization] = DEFAULT_UNIFORM,
                 prior: Optional[Parameter] = DEFAULT_SCALED_GAUSSIAN_MIXTURE) -> None:
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.initialization = initialization

        self.weight = Gaussian(Size((out_features, in_features)), self.initialization)
        self.weight_prior = prior
        if bias:
            self.bias = Gaussian(Size((out_features,)), self.initialization)
            self.bias_prior = prior
        else:
            self.bias = NoneParameter()
            self.bias_prior = NoneParameter()
        # extension (default None = the reference's plain affine map): "gelu" applies the exact GELU inside the
        # layer, fused into the tensor-core epilogue (forward) and into the bias-gradient pass (backward)
        self.activation: Optional[str] = None
        # one-shot hand-off set by a fused consumer of this layer's output (layers/fused.py): a list its backward
        # fills with the row sums of the output gradient, so this layer's backward skips its own bias-gradient pass
        self._bias_grad_box: Optional[list] = None
        self._init_scalars()

    def forward(self, input: Tensor) -> Tensor:
        S = runtime.get_mc_samples()
        kl_grad = self._kl_grad()
        has_bias = isinstance(self.bias, Gaussian)
        w_prior = prior_spec_of(self.weight_prior)
        b_prior = prior_spec_of(self.bias_prior) if has_bias else ops.PriorSpec()
        pre, self._presampled = self._presampled, None  # a draw serves exactly one forward
        if pre is not None and pre[0] != S:
            pre = None
        if pre is not None:
            spec = ops.LinearSpec(S=S, gemm_dtype=self._gemm_dtype(), kl_grad=kl_grad, w_prior=w_prior, b_prior=b_prior,
                                  w_stream=pre[5], b_stream=pre[6], presampled=pre[1:5],
                                  activation=self.activation)
        else:
            spec = ops.LinearSpec(S=S, gemm_dtype=self._gemm_dtype(), kl_grad=kl_grad, w_prior=w_prior,
                                  b_prior=b_prior, w_stream=self.weight.next_stream(S),
                                  b_stream=self.bias.next_stream(S) if has_bias else ops.StreamSpec(),
                                  activation=self.activation)
        spec.bias_grad_box, self._bias_grad_box = self._bias_grad_box, None
        if runtime.grad_sinks_enabled() and torch.is_grad_enabled() and input.requires_grad:
            spec.sink = runtime.sink_for(input, create=True)
        links = runtime.gelu_links_enabled() and torch.is_grad_enabled()
        if links:
            spec.gelu_in = getattr(input, "_bf_gelu_link", None)  # input = gelu(z) of a fused Bayesian Linear?
            if self.activation == "gelu":
                spec.gelu_out = []
        self._last_streams = (spec.w_stream, spec.b_stream)  # identity of this forward's eps draw (tests, debugging)
        y, logq, logp = ops.BayesLinear.apply(
            input, self.weight.mu, self.weight.rho,
            self.bias.mu if has_bias else None, self.bias.rho if has_bias else None,
            w_prior.mu, w_prior.rho, b_prior.mu, b_prior.rho, spec)
        self._publish(logq, logp, S, kl_grad, means=pre[7] if pre is not None and len(pre) > 7 else None)
        if spec.gelu_out:  # fused GELU ran: let the consumer of y fold gelu'(z) into its dgrad (runtime.GeluLink)
            link = spec.gelu_out[0]
            y._bf_gelu_link = link
            if y.requires_grad:
                y.register_hook(link.check)
        return y

    @classmethod
    def from_frequentist(cls, linear: nn.Module, initialization: Optional[Initialization] = DEFAULT_UNIFORM,
                         prior: Optional[Parameter] = DEFAULT_SCALED_GAUSSIAN_MIXTURE, delta: float = None,
                         freeze: bool = False) -> "Linear":
        # `initialization` is accepted and not forwarded, like the reference (linear.py:137, quirk Q3)
        has_bias = linear.bias is not None
        baye = cls(linear.in_features, linear.out_features, has_bias, prior=prior)
        if delta is not None:
            baye.weight_prior = moped_(baye.weight, linear.weight, delta, freeze)
            if has_bias:
                baye.bias_prior = moped_(baye.bias, linear.bias, delta, freeze)
        return baye

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={isinstance(self.bias, Gaussian)}"
