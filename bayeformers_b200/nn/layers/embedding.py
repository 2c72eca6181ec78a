"""Bayesian Embedding (absent from the reference snapshot).

Specified by analogy with bnn.Linear: the whole table is a `Gaussian`, every
forward samples the whole table and reduces log q / log p over it (one fused
kernel pass), then looks rows up.  MOPED conversion as for Linear.
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
        out_dtype = torch.float32
        w, logq, logp = ops.SampleKL.apply(self.weight.mu, self.weight.rho, prior.mu, prior.rho, prior,
                                           self.weight.next_stream(S), S, out_dtype, kl_grad)
        self._publish(logq, logp, S, kl_grad)
        kw = dict(padding_idx=self.padding_idx, max_norm=self.max_norm, norm_type=self.norm_type,
                  scale_grad_by_freq=self.scale_grad_by_freq)
        if S == 1:
            return F.embedding(input, w[0], **kw)
        rows = runtime.get_folded_rows()
        if input.shape[0] == 1 and rows is not None and rows % S == 0:
            # broadcast ids (e.g. HF position_ids [1, T]): every folded row looks its ids up in ITS sample's table
            input = input.expand(rows, *input.shape[1:])
        if input.shape[0] % S != 0:
            raise ValueError(f"leading dimension {input.shape[0]} is not a multiple of mc_samples={S}")
        chunks = input.view(S, input.shape[0] // S, *input.shape[1:])
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
