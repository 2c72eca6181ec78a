"""Bayesian Embedding (absent from the reference snapshot).

Specified by analogy with bnn.Linear: the whole table is a `Gaussian`, every
forward draws a sample of the whole table, reduces log q / log p over ALL of it
and looks rows up (the reference's `Gaussian.sample`, gaussian.py:90-101,
composed with `F.embedding`).  MOPED conversion as for Linear.

The S sampled tables are never materialised: eps is a pure function of (seed,
step, tensor, sample, element), so `bf_embedding_fwd` samples exactly the rows
the ids name while a weight-less sample+KL pass reduces the two log-probs over
the whole table with the same stream; the backward is row-sparse and
deterministic (`bf_embedding_bwd`).  Output dtype = the layer's GEMM dtype
(fp32, or bf16 in tensor-core mode).  `max_norm` / `scale_grad_by_freq` fall back
to torch's semantics on a materialised table.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Size, Tensor

from ... import ops, runtime
from ..parameters.base import Parameter
from ..parameters.gaussian import DEFAULT_SCALED_GAUSSIAN_MIXTURE, Gaussian, prior_spec_of
from ..parameters.initializations import DEFAULT_UNIFORM, Initialization
from .common import BayesianLayer, moped_


class Embedding(BayesianLayer):
    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = None,
                 max_norm: Optional[float] = None, norm_type: float = 2.0, scale_grad_by_freq: bool = False,
                 sparse: bool = False, initialization: Optional[Initialization] = DEFAULT_UNIFORM,
                 prior: Optional[Parameter] = DEFAULT_SCALED_GAUSSIAN_MIXTURE) -> None:
        super().__init__()
        if sparse:
            raise ValueError("sparse gradients are not supported: the variational backward needs the dense table gradient")
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.padding_idx, self.max_norm, self.norm_type = padding_idx, max_norm, norm_type
        self.scale_grad_by_freq = scale_grad_by_freq
        self.initialization = initialization
        self.weight = Gaussian(Size((num_embeddings, embedding_dim)), self.initialization)
        self.weight_prior = prior
        self._init_scalars()

    def forward(self, input: Tensor) -> Tensor:
        S = runtime.get_mc_samples()
        kl_grad = self._kl_grad()
        prior = prior_spec_of(self.weight_prior)
        if S > 1:
            rows = runtime.get_folded_rows()
            if input.shape[0] == 1 and rows is not None and rows % S == 0:
                # broadcast ids (e.g. HF position_ids [1, T]): every folded row looks its ids up in ITS sample's table
                input = input.expand(rows, *input.shape[1:])
            if input.shape[0] % S != 0:
                raise ValueError(f"leading dimension {input.shape[0]} is not a multiple of mc_samples={S}")
        pre, self._presampled = self._presampled, None  # a draw serves exactly one forward
        if pre is not None and pre[0] != S:
            pre = None
        fused = (self.max_norm is None and not self.scale_grad_by_freq and ops.embedding_supported(self.embedding_dim)
                 and input.dim() >= 1)
        if fused:
            # rows are sampled on lookup; log q / log p still cover the whole table (module docstring)
            stream = pre[5] if pre is not None else self.weight.next_stream(S)
            spec = ops.EmbeddingSpec(S=S, kl_grad=kl_grad, prior=prior, stream=stream, out_dtype=runtime.activation_dtype(self._gemm_dtype()),
                                     padding_idx=self.padding_idx,
                                     presampled=None if pre is None else (pre[3], pre[4]))
            self._last_streams = (stream, None)
            out, logq, logp = ops.EmbeddingFn.apply(input, self.weight.mu, self.weight.rho, prior.mu, prior.rho, spec)
            self._publish(logq, logp, S, kl_grad, means=pre[7] if pre is not None else None)
            return out
        # max_norm / scale_grad_by_freq: torch semantics on a materialised table (not the hot path)
        w, logq, logp = ops.SampleKL.apply(self.weight.mu, self.weight.rho, prior.mu, prior.rho, prior,
                                           self.weight.next_stream(S), S, torch.float32, kl_grad)
        self._publish(logq, logp, S, kl_grad)
        kw = dict(padding_idx=self.padding_idx, max_norm=self.max_norm, norm_type=self.norm_type,
                  scale_grad_by_freq=self.scale_grad_by_freq)
        if S == 1:
            return F.embedding(input, w[0], **kw)
        chunks = input.reshape(S, input.shape[0] // S, *input.shape[1:])
        return torch.cat([F.embedding(chunks[s], w[s], **kw) for s in range(S)], dim=0)

    @classmethod
    def from_frequentist(cls, emb: nn.Module, initialization: Optional[Initialization] = DEFAULT_UNIFORM,
                         prior: Optional[Parameter] = DEFAULT_SCALED_GAUSSIAN_MIXTURE, delta: float = None,
                         freeze: bool = False) -> "Embedding":
        baye = cls(emb.num_embeddings, emb.embedding_dim, emb.padding_idx, emb.max_norm, emb.norm_type,
                   emb.scale_grad_by_freq, False, prior=prior)
        if delta is not None:
            baye.weight_prior = moped_(baye.weight, emb.weight, delta, freeze)
        return baye
