"""Shared pieces of the Bayesian layers: MOPED initialisation and bookkeeping
of the two per-layer scalars."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ... import runtime
from ..parameters.gaussian import Gaussian


def moped_(posterior: Gaussian, source: torch.Tensor, delta: float, freeze: bool) -> Gaussian:
    """MOPED (Krishnan et al. 2020) initialisation of `posterior` from the
    pretrained tensor `source`, and the matching Gaussian prior.  Kept as the
    reference's literal torch op sequence -- NOT expm1, NOT a custom kernel --
    because bit-exact mu/rho is the contract
    (/root/reference/bayeformers/nn/layers/linear.py:140-150):
        mu   <- source (same storage, no copy)
        rho  <- log(exp(delta*|source|) - 1), with -inf -> 0
        prior = Gaussian(mu = source (same storage), rho = 1)
    Returns the prior."""
    posterior.mu.data = source.data
    posterior.rho.data = torch.log(torch.exp(delta * torch.abs(source.data)) - 1.0)
    posterior.rho.data[posterior.rho.data == float("-inf")] = 0.0
    posterior.mu.requires_grad = not freeze

    prior = Gaussian(posterior.mu.size())
    prior.mu.data = source.data
    prior.rho.data = torch.ones_like(source)
    return prior


class BayesianLayer(nn.Module):
    """Holds `log_prior` / `log_variational_posterior` exactly as the reference
    does (0-dim requires_grad=False Parameters that appear in state_dict,
    linear.py:80-81, quirk Q2) plus the per-layer switches of the extensions."""

    def _init_scalars(self) -> None:
        self.register_parameter("log_prior", nn.Parameter(torch.tensor(0.), requires_grad=False))
        self.register_parameter("log_variational_posterior", nn.Parameter(torch.tensor(0.), requires_grad=False))
        self.kl_grad: Optional[bool] = None       # None -> runtime.get_kl_grad()
        self.gemm_dtype: Optional[torch.dtype] = None  # None -> runtime.get_gemm_dtype()
        # grad-carrying versions of the two scalars when kl_grad is on
        self.live_log_prior = None
        self.live_log_variational_posterior = None
        # per-sample [S] values of the last folded forward (None when S == 1); plain attributes, never in state_dict
        self.log_prior_samples = None
        self.log_variational_posterior_samples = None
        self._presampled = None  # one forward's draw made by presample.Presampler (consumed by forward)

    def _kl_grad(self) -> bool:
        return runtime.get_kl_grad() if self.kl_grad is None else self.kl_grad

    def _gemm_dtype(self) -> torch.dtype:
        return runtime.get_gemm_dtype() if self.gemm_dtype is None else self.gemm_dtype

    def _publish(self, logq: torch.Tensor, logp: torch.Tensor, S: int, kl_grad: bool, means=None) -> None:
        """Store the last forward's scalars.  The two registered Parameters stay 0-dim whatever S is (so a
        state_dict saved after a folded forward has the reference's shapes, linear.py:80-81): the value itself for
        S == 1 (reference behaviour), the mean over the S samples under mc_samples(S) -- what downstream code takes
        of them anyway (examples/bert_glue.py:70-71).  The per-sample [S] values live in `*_samples` (and, grad
        carrying, in `live_*` when kl_grad is on); `bnn.Model.log_prior()` sums those.
        `means`: optional precomputed (mean log q, mean log p) 0-dim tensors (multi-tensor sampling computes them for
        all layers at once)."""
        if S == 1:
            lq, lp = logq[0], logp[0]
            self.log_prior.data = lp.detach()
            self.log_variational_posterior.data = lq.detach()
            self.log_prior_samples = self.log_variational_posterior_samples = None
        else:
            lq, lp = logq, logp
            mq, mp = means if means is not None else (logq.detach().mean(), logp.detach().mean())
            self.log_prior.data = mp
            self.log_variational_posterior.data = mq
            self.log_prior_samples = lp.detach()
            self.log_variational_posterior_samples = lq.detach()
        self.live_log_prior = lp if kl_grad else None
        self.live_log_variational_posterior = lq if kl_grad else None

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # checkpoints written by round-1 builds may hold [S]-shaped scalars: fold them to the reference's 0-dim form
        for name in ("log_prior", "log_variational_posterior"):
            v = state_dict.get(prefix + name)
            if torch.is_tensor(v) and v.dim() > 0:
                state_dict[prefix + name] = v.float().mean()
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
