"""Gaussian variational parameter and the scale-mixture prior.

API parity with /root/reference/bayeformers/nn/parameters/gaussian.py:22-177
(same constructor arguments, attribute and state_dict names).  What changed is
where the arithmetic runs: `sample()` is one fused CUDA pass (Philox eps in
registers, w = mu + softplus(rho)*eps, log q reduced on the fly) instead of
~25 separate torch kernels, and backward regenerates eps from the counter.

`log_prob(input)` on an *arbitrary* tensor is kept as the reference's formula
in torch ops: it is the compatibility surface (and the differentiable oracle
for `kl_grad=True`), not the hot path -- layers never call it.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Size, Tensor
from torch.distributions.normal import Normal

from ... import ops, runtime
from ..._lib import BF_PRIOR_GAUSSIAN, BF_PRIOR_MIXTURE, BF_PRIOR_NONE
from .base import Parameter, parameter
from .initializations import DEFAULT_UNIFORM, Initialization

_LOG_SQRT_2PI = np.log(np.sqrt(2 * np.pi))


def _fixed(value: float) -> nn.Parameter:
    return nn.Parameter(torch.tensor(value).float(), requires_grad=False)


class Gaussian(Parameter):
    """w ~ N(mu, softplus(rho)^2), reparametrised as w = mu + softplus(rho) * eps."""

    def __init__(self, size: Size, initialization: Optional[Initialization] = DEFAULT_UNIFORM,
                 dtype: Optional[torch.dtype] = torch.float32) -> None:
        super().__init__()
        self.size, self.dtype = size, dtype
        self.initialization = initialization
        self.mu = parameter(self.size, dtype=self.dtype)
        self.rho = parameter(self.size, dtype=self.dtype)
        self.register_parameter("zero", _fixed(0.0))
        self.register_parameter("one", _fixed(1.0))
        # kept for API parity and as the eps-injection point: replacing this
        # attribute by any object with `.sample(size)` overrides the Philox
        # stream (the FixedEps trick the parity tests use on the reference too)
        self.normal = Normal(self.zero, self.one)
        self.tensor_id = runtime.next_tensor_id()
        self.step = 0
        self.last_log_prob = None
        self.reset_parameters()

    def reset_parameters(self) -> None:
        self.mu, self.rho = self.initialization(self.mu, self.rho)

    def __deepcopy__(self, memo):
        # a copy is a NEW variational tensor: it must not share the eps stream of the original (cloned layers of
        # nn.TransformerEncoder, EMA copies, ... would otherwise draw identical eps every step)
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        from copy import deepcopy
        for k, v in self.__dict__.items():
            new.__dict__[k] = deepcopy(v, memo)
        new.normal = Normal(new.zero, new.one) if isinstance(self.normal, Normal) else self.normal
        new.tensor_id = runtime.next_tensor_id()
        return new

    @property
    def sigma(self) -> Tensor:
        return F.softplus(self.rho)

    # ---- sigma cache (extension) -------------------------------------------
    # sigma = softplus(rho) changes only when rho does.  `bf.optim.ClipAdamW(..., model=...)` writes softplus of the
    # updated rho next to its update, and the multi-tensor sampling kernel then reads this tensor instead of rho (no
    # exp / log1p per element).  The cache counts as valid only while (storage, version, device) of rho are what they were
    # when it was written: any torch-side change of rho (load_state_dict, copy_, another optimizer, .to()) invalidates it
    # and the kernels go back to computing softplus themselves.
    def sigma_cache(self) -> Optional[Tensor]:
        """The cached softplus(rho) when it is valid, else None."""
        sig = getattr(self, "_sigma", None)
        if sig is not None and self._sigma_key == (self.rho.data_ptr(), self.rho._version, self.rho.device):
            return sig
        return None

    def refresh_sigma_cache(self, written_by_kernel: bool = False) -> Tensor:
        """(Re)fill the cache from rho -- or, when a kernel has just written it, only mark it current."""
        sig = getattr(self, "_sigma", None)
        if sig is None or sig.shape != self.rho.shape or sig.device != self.rho.device:
            sig = torch.empty_like(self.rho, dtype=torch.float32)
            written_by_kernel = False
        if not written_by_kernel:
            from ... import _lib
            with ops.on_device(self.rho.device):
                _lib.check(_lib.load().bf_softplus_fwd(self.rho.data_ptr(), sig.data_ptr(), self.rho.numel(),
                                                       ops._stream(self.rho.device)), "bf_softplus_fwd")
        object.__setattr__(self, "_sigma", sig)  # a plain attribute: never a Parameter / buffer / state_dict entry
        self._sigma_key = (self.rho.data_ptr(), self.rho._version, self.rho.device)
        return sig

    # ---- eps stream -------------------------------------------------------
    def next_stream(self, S: int = 1) -> ops.StreamSpec:
        """Identity of this call's eps draw; advances the per-tensor step."""
        eps = None
        if not isinstance(self.normal, Normal):
            eps = torch.stack([torch.as_tensor(self.normal.sample(self.size)) for _ in range(S)])
        spec = ops.StreamSpec(seed=runtime.seed(), tensor_id=self.tensor_id, step=self.step & 0xFFFFFFFF, eps=eps)
        self.step += 1
        return spec

    def prior_spec(self) -> ops.PriorSpec:
        """This distribution used as a prior over another tensor (MOPED).  When
        rho is one constant (MOPED sets it to 1 everywhere, linear.py:149) the
        kernels take sigma_p as a scalar instead of reading the array; the check
        is cached per (storage, version), i.e. one device sync per layer."""
        key = (self.rho.data_ptr(), self.rho._version, self.rho.device)
        if getattr(self, "_const_key", None) != key:
            flat = self.rho.detach().reshape(-1)
            self._const_sigma = None
            if flat.numel() > 0 and bool((flat == flat[0]).all()):
                self._const_sigma = float(F.softplus(flat[0]))
            self._const_key = key
        if self._const_sigma is not None:
            return ops.PriorSpec(kind=BF_PRIOR_GAUSSIAN, sigma1=self._const_sigma, mu=self.mu, rho=None)
        return ops.PriorSpec(kind=BF_PRIOR_GAUSSIAN, mu=self.mu, rho=self.rho)

    def sample(self, mc_samples: Optional[int] = None) -> Tensor:
        """One fused pass; differentiable w.r.t. (mu, rho).  Shape `size`
        (or [S, *size] when mc_samples is given).  log q of the draw is left in
        `self.last_log_prob` ([S])."""
        S = 1 if mc_samples is None else int(mc_samples)
        w, logq, _ = ops.SampleKL.apply(self.mu, self.rho, None, None, ops.PriorSpec(kind=BF_PRIOR_NONE),
                                        self.next_stream(S), S, torch.float32, runtime.get_kl_grad())
        self.last_log_prob = logq
        return w[0] if mc_samples is None else w

    def log_prob(self, input: Tensor) -> Tensor:
        sigma = self.sigma
        return (-_LOG_SQRT_2PI - torch.log(sigma) - ((input - self.mu) ** 2) / (2 * self.sigma ** 2)).sum()


class ScaledGaussianMixture(Parameter):
    """pi * N(0, sigma1^2) + (1 - pi) * N(0, sigma2^2), prior only (sample() is a
    stub returning 0.0 exactly like the reference, gaussian.py:152-158)."""

    def __init__(self, pi: float, sigma1: float, sigma2: float) -> None:
        super().__init__()
        self.register_parameter("pi", _fixed(pi))
        self.register_parameter("sigma1", _fixed(sigma1))
        self.register_parameter("sigma2", _fixed(sigma2))
        self.register_parameter("zero", _fixed(0.0))
        self.gaussian1 = Normal(self.zero, self.sigma1)
        self.gaussian2 = Normal(self.zero, self.sigma2)
        self._scalars = (float(self.pi), float(self.sigma1), float(self.sigma2))

    def prior_spec(self) -> ops.PriorSpec:
        # scalars travel as kernel arguments (quirk Q11); cached so no .item() sync per forward
        pi, s1, s2 = self._scalars
        return ops.PriorSpec(kind=BF_PRIOR_MIXTURE, pi=pi, sigma1=s1, sigma2=s2)

    def refresh(self) -> None:
        """Re-read the scalars after editing pi/sigma1/sigma2 in place."""
        self._scalars = (float(self.pi), float(self.sigma1), float(self.sigma2))

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.refresh()

    def sample(self) -> float:
        return 0.0

    def log_prob(self, input: Tensor) -> Tensor:
        # Normal(...) is rebuilt from the current parameters so device moves are honoured
        p1 = torch.exp(Normal(self.zero, self.sigma1).log_prob(input))
        p2 = torch.exp(Normal(self.zero, self.sigma2).log_prob(input))
        return torch.log(self.pi * p1 + (1.0 - self.pi) * p2).sum()


DEFAULT_SCALED_GAUSSIAN_MIXTURE = ScaledGaussianMixture(0.5, np.exp(-0), np.exp(-6))


def prior_spec_of(prior) -> ops.PriorSpec:
    """PriorSpec of whatever object a layer holds as `*_prior`."""
    if hasattr(prior, "prior_spec"):
        return prior.prior_spec()
    from .base import NoneParameter
    if prior is None or isinstance(prior, NoneParameter):
        return ops.PriorSpec(kind=BF_PRIOR_NONE)
    raise TypeError(
        f"prior of type {type(prior).__name__} has no CUDA kernel; supported: Gaussian, ScaledGaussianMixture")
