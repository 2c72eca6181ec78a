"""CPU oracle for the variational-layer hot path -- TEST INFRASTRUCTURE ONLY.

Every function restates, in plain functional form, what one piece of the
reference computes, using the same fp32 torch-CPU operator sequence (the
reference's arithmetic *is* torch's: SURVEY.md section 8c), and cites the
reference file:line it follows (paths relative to /root/reference).  Float64
numpy closed forms sit next to them for the known-answer tests and for the
KL-gradient checks, where fp32 autograd cancels badly.

Pinned by `tests/golden/*.npz` (generated from the unmodified reference by
`tests/golden/make_golden.py`).  Not imported by the product package.
"""
from __future__ import annotations

import copy
import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# numpy float64 constant folded into fp32 tensor arithmetic by the reference
# (bayeformers/nn/parameters/gaussian.py:113).
LOG_SQRT_2PI = float(np.log(np.sqrt(2 * np.pi)))

# default scale-mixture prior (bayeformers/nn/parameters/gaussian.py:175-177)
DEFAULT_PI = 0.5
DEFAULT_SIGMA1 = float(np.exp(-0))
DEFAULT_SIGMA2 = float(np.exp(-6))


# --------------------------------------------------------------------------- #
# fp32 torch restatement (bit-faithful to the reference's operator sequence)   #
# --------------------------------------------------------------------------- #
def sigma_of(rho: torch.Tensor) -> torch.Tensor:
    """sigma = softplus(rho), beta=1, threshold=20.
    Ref: bayeformers/nn/parameters/gaussian.py:81-88."""
    return F.softplus(rho)


def gaussian_sample(mu: torch.Tensor, rho: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """w = mu + eps * sigma (two roundings: product, then sum).
    Ref: bayeformers/nn/parameters/gaussian.py:100-101."""
    return mu + eps * sigma_of(rho)


def gaussian_log_prob(w: torch.Tensor, mu: torch.Tensor, rho: torch.Tensor) -> torch.Tensor:
    """sum(-log sqrt(2pi) - log sigma - (w-mu)^2 / (2 sigma^2)).
    Used for log q (posterior) and for log p under a MOPED Gaussian prior.
    Ref: bayeformers/nn/parameters/gaussian.py:103-116."""
    s = sigma_of(rho)
    s_again = sigma_of(rho)  # the reference evaluates the property twice
    return (-LOG_SQRT_2PI - torch.log(s) - ((w - mu) ** 2) / (2 * s_again ** 2)).sum()


def _normal_log_prob(x: torch.Tensor, loc: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """torch.distributions.Normal(loc, scale).log_prob(x) with tensor scale
    (third-party: torch/distributions/normal.py, torch 2.11):
    -((x-loc)^2)/(2 var) - log(scale) - log(sqrt(2 pi))."""
    var = scale ** 2
    return -((x - loc) ** 2) / (2 * var) - scale.log() - math.log(math.sqrt(2 * math.pi))


def mixture_log_prob(w: torch.Tensor, pi: float = DEFAULT_PI, sigma1: float = DEFAULT_SIGMA1,
                     sigma2: float = DEFAULT_SIGMA2) -> torch.Tensor:
    """sum log(pi * exp(N(w;0,s1)) + (1-pi) * exp(N(w;0,s2))) -- explicit
    exp-then-log (no logsumexp), scalars held as fp32 0-dim tensors.
    Ref: bayeformers/nn/parameters/gaussian.py:140-150 (ctor), :160-171."""
    t_pi = torch.tensor(pi).float()
    t_s1 = torch.tensor(sigma1).float()
    t_s2 = torch.tensor(sigma2).float()
    zero = torch.tensor(0.0).float()
    p1 = torch.exp(_normal_log_prob(w, zero, t_s1))
    p2 = torch.exp(_normal_log_prob(w, zero, t_s2))
    return torch.log(t_pi * p1 + (1.0 - t_pi) * p2).sum()


def mixture_log_prob_elementwise(w: torch.Tensor, pi: float = DEFAULT_PI, sigma1: float = DEFAULT_SIGMA1,
                                 sigma2: float = DEFAULT_SIGMA2) -> torch.Tensor:
    """Same as `mixture_log_prob` without the final sum (for the KATs)."""
    t_pi = torch.tensor(pi).float()
    t_s1 = torch.tensor(sigma1).float()
    t_s2 = torch.tensor(sigma2).float()
    zero = torch.tensor(0.0).float()
    p1 = torch.exp(_normal_log_prob(w, zero, t_s1))
    p2 = torch.exp(_normal_log_prob(w, zero, t_s2))
    return torch.log(t_pi * p1 + (1.0 - t_pi) * p2)


def moped_rho(weight: torch.Tensor, delta: float) -> torch.Tensor:
    """MOPED posterior scale: rho = log(exp(delta*|w|) - 1), -inf -> 0.
    NOT expm1 -- the reference's literal op sequence is what "bit-exact" means.
    Ref: bayeformers/nn/layers/linear.py:141-144 (weight), :154-157 (bias)."""
    rho = torch.log(torch.exp(delta * torch.abs(weight)) - 1.0)
    rho[rho == float("-inf")] = 0.0
    return rho


def uniform_init(mu: torch.Tensor, rho: torch.Tensor, mu_range=(-0.2, 0.2), rho_range=(-5.0, -4.0)):
    """Default initialisation: mu ~ U(mu_range) first, then rho ~ U(rho_range),
    both from torch's global generator.
    Ref: bayeformers/nn/parameters/initializations.py:54-55, default :60."""
    mu.uniform_(*mu_range)
    rho.uniform_(*rho_range)
    return mu, rho


# prior descriptions used by the functional layer oracle
def default_mixture_prior() -> Dict:
    return {"kind": "mixture", "pi": DEFAULT_PI, "sigma1": DEFAULT_SIGMA1, "sigma2": DEFAULT_SIGMA2}


def gaussian_prior(mu_p: torch.Tensor, rho_p: torch.Tensor) -> Dict:
    return {"kind": "gaussian", "mu": mu_p, "rho": rho_p}


def prior_log_prob(w: torch.Tensor, prior: Dict) -> torch.Tensor:
    if prior["kind"] == "mixture":
        return mixture_log_prob(w, prior["pi"], prior["sigma1"], prior["sigma2"])
    if prior["kind"] == "gaussian":
        return gaussian_log_prob(w, prior["mu"], prior["rho"])
    raise ValueError(prior["kind"])


def linear_forward(x: torch.Tensor, w_mu, w_rho, b_mu, b_rho, eps_w, eps_b,
                   w_prior: Dict, b_prior: Optional[Dict]):
    """One forward of the Bayesian Linear for ONE Monte-Carlo sample.
    Returns (y, log_prior, log_variational_posterior, W, b); the two scalars
    are detached exactly like the reference's `.data =` assignment.
    Ref: bayeformers/nn/layers/linear.py:83-104."""
    W = gaussian_sample(w_mu, w_rho, eps_w)
    b = gaussian_sample(b_mu, b_rho, eps_b) if b_mu is not None else None
    with torch.no_grad():
        lp = prior_log_prob(W, w_prior)
        lq = gaussian_log_prob(W, w_mu, w_rho)
        if b is not None:
            lp = lp + prior_log_prob(b, b_prior)
            lq = lq + gaussian_log_prob(b, b_mu, b_rho)
    y = F.linear(x, W, b)
    return y, lp, lq, W, b


def elbo_terms_with_grad(mu, rho, eps, prior: Dict):
    """log q(w) and log p(w) WITH autograd attached (what the reference's own
    Gaussian.log_prob / ScaledGaussianMixture.log_prob give when called
    directly, i.e. without Linear.forward's `.data` detach).  This is the
    `kl_grad=True` oracle of SURVEY.md section 8c."""
    w = gaussian_sample(mu, rho, eps)
    return gaussian_log_prob(w, mu, rho), prior_log_prob(w, prior), w


# --------------------------------------------------------------------------- #
# float64 closed forms (numpy)                                                 #
# --------------------------------------------------------------------------- #
def softplus_f64(rho: np.ndarray) -> np.ndarray:
    rho = np.asarray(rho, dtype=np.float64)
    return np.where(rho > 20.0, rho, np.log1p(np.exp(np.minimum(rho, 20.0))))


def gaussian_log_prob_f64(w, mu, rho) -> float:
    w, mu = np.asarray(w, np.float64), np.asarray(mu, np.float64)
    s = softplus_f64(rho)
    return float(np.sum(-0.5 * np.log(2 * np.pi) - np.log(s) - (w - mu) ** 2 / (2 * s ** 2)))


def mixture_log_prob_elementwise_f64(w, pi=DEFAULT_PI, sigma1=DEFAULT_SIGMA1, sigma2=DEFAULT_SIGMA2):
    w = np.asarray(w, np.float64)
    c = 0.5 * np.log(2 * np.pi)
    n1 = -w ** 2 / (2 * sigma1 ** 2) - np.log(sigma1) - c
    n2 = -w ** 2 / (2 * sigma2 ** 2) - np.log(sigma2) - c
    return np.logaddexp(np.log(pi) + n1, np.log1p(-pi) + n2)


def dlogp_dw_f64(w, prior: Dict) -> np.ndarray:
    """d log p(w) / dw, closed form."""
    w = np.asarray(w, np.float64)
    if prior["kind"] == "gaussian":
        mu_p = np.asarray(prior["mu"], np.float64)
        s_p = softplus_f64(np.asarray(prior["rho"], np.float64))
        return -(w - mu_p) / s_p ** 2
    pi, s1, s2 = prior["pi"], prior["sigma1"], prior["sigma2"]
    # responsibilities in a numerically safe form
    l1 = np.log(pi) - np.log(s1) - w ** 2 / (2 * s1 ** 2)
    l2 = np.log1p(-pi) - np.log(s2) - w ** 2 / (2 * s2 ** 2)
    m = np.maximum(l1, l2)
    a1, a2 = np.exp(l1 - m), np.exp(l2 - m)
    return -w * (a1 / s1 ** 2 + a2 / s2 ** 2) / (a1 + a2)


def kl_grads_f64(mu, rho, eps, prior: Dict, g_logq: float, g_logp: float):
    """Gradient of g_logq*log q(w) + g_logp*log p(w), w = mu + softplus(rho)*eps,
    w.r.t. (mu, rho) in float64.  d log q/d mu cancels exactly; d log q/d sigma =
    -1/sigma; the prior contributes p'(w) to mu and p'(w)*eps to sigma."""
    mu, rho, eps = (np.asarray(a, np.float64) for a in (mu, rho, eps))
    s = softplus_f64(rho)
    w = mu + s * eps
    dp = dlogp_dw_f64(w, prior)
    sig = 1.0 / (1.0 + np.exp(-rho))  # d softplus / d rho
    g_mu = g_logp * dp
    g_rho = (g_logq * (-1.0 / s) + g_logp * dp * eps) * sig
    return g_mu, g_rho


# --------------------------------------------------------------------------- #
# module-shaped oracle (lets the restatement sit inside a host model such as   #
# HF BERT for the CPU baseline / the `--impl reference` arm of bench.py)       #
# --------------------------------------------------------------------------- #
class EpsSource:
    """Where eps comes from.  Default: the reference's own draw,
    Normal(0,1).sample(size) == torch.normal(zeros, ones) under no_grad
    (bayeformers/nn/parameters/gaussian.py:71,100).  Tests replace `draw`
    with a queue of preset tensors (the FixedEps trick of SURVEY.md 8c)."""

    def __init__(self, preset: Optional[List[torch.Tensor]] = None):
        self.preset = list(preset) if preset is not None else None
        self._zero = torch.tensor(0.0)
        self._one = torch.tensor(1.0)

    def draw(self, size) -> torch.Tensor:
        if self.preset is not None:
            e = self.preset.pop(0)
            assert tuple(e.shape) == tuple(size), (tuple(e.shape), tuple(size))
            return e
        with torch.no_grad():
            return torch.normal(self._zero.expand(size), self._one.expand(size))


class OracleLinear(nn.Module):
    """Module wrapper around `linear_forward` (one MC sample per call).
    Ref: bayeformers/nn/layers/linear.py:24-104."""

    def __init__(self, w_mu, w_rho, b_mu, b_rho, w_prior, b_prior, eps: EpsSource, mu_trainable=True):
        super().__init__()
        self.w_mu = nn.Parameter(w_mu, requires_grad=mu_trainable)
        self.w_rho = nn.Parameter(w_rho)
        self.b_mu = nn.Parameter(b_mu, requires_grad=mu_trainable) if b_mu is not None else None
        self.b_rho = nn.Parameter(b_rho) if b_rho is not None else None
        self.w_prior, self.b_prior = w_prior, b_prior
        self.eps = eps
        self.log_prior = torch.tensor(0.0)
        self.log_variational_posterior = torch.tensor(0.0)

    def forward(self, x):
        eps_w = self.eps.draw(self.w_mu.shape)
        eps_b = self.eps.draw(self.b_mu.shape) if self.b_mu is not None else None
        y, lp, lq, _, _ = linear_forward(x, self.w_mu, self.w_rho, self.b_mu, self.b_rho,
                                         eps_w, eps_b, self.w_prior, self.b_prior)
        self.log_prior, self.log_variational_posterior = lp, lq
        return y


class _OracleVariational(nn.Module):
    """Shared by the two modules the reference snapshot lacks: they are SPECIFIED as the
    reference's Gaussian.sample / .log_prob (gaussian.py:90-116) composed with the torch functional op, with the
    log-prob bookkeeping of Linear.forward (linear.py:97-102: weight then bias, detached)."""

    def _draw(self, names):
        ws, lp, lq = [], 0.0, 0.0
        for n in names:
            mu, rho, prior = getattr(self, n + "_mu"), getattr(self, n + "_rho"), getattr(self, n + "_prior")
            if mu is None:
                ws.append(None)
                continue
            w = gaussian_sample(mu, rho, self.eps.draw(mu.shape))
            with torch.no_grad():
                lp = lp + prior_log_prob(w, prior)
                lq = lq + gaussian_log_prob(w, mu, rho)
            ws.append(w)
        self.log_prior, self.log_variational_posterior = lp, lq
        return ws


class OracleEmbedding(_OracleVariational):
    """Sample the WHOLE table, log-probs over the whole table, then F.embedding."""

    def __init__(self, w_mu, w_rho, w_prior, eps: EpsSource, padding_idx=None, mu_trainable=True):
        super().__init__()
        self.w_mu = nn.Parameter(w_mu, requires_grad=mu_trainable)
        self.w_rho = nn.Parameter(w_rho)
        self.w_prior, self.eps, self.padding_idx = w_prior, eps, padding_idx
        self.log_prior = self.log_variational_posterior = torch.tensor(0.0)

    def forward(self, ids):
        (W,) = self._draw(["w"])
        return F.embedding(ids, W, padding_idx=self.padding_idx)


class OracleLayerNorm(_OracleVariational):
    """Sample gamma and beta, then F.layer_norm."""

    def __init__(self, shape, ln_eps, w_mu, w_rho, b_mu, b_rho, w_prior, b_prior, eps: EpsSource, mu_trainable=True):
        super().__init__()
        self.shape, self.ln_eps = tuple(shape), ln_eps
        self.w_mu = nn.Parameter(w_mu, requires_grad=mu_trainable)
        self.w_rho = nn.Parameter(w_rho)
        self.b_mu = nn.Parameter(b_mu, requires_grad=mu_trainable) if b_mu is not None else None
        self.b_rho = nn.Parameter(b_rho) if b_rho is not None else None
        self.w_prior, self.b_prior, self.eps = w_prior, b_prior, eps
        self.log_prior = self.log_variational_posterior = torch.tensor(0.0)

    def forward(self, x):
        W, b = self._draw(["w", "b"])
        return F.layer_norm(x, self.shape, W, b, self.ln_eps)


def _moped_pair(src: torch.Tensor, delta: Optional[float]):
    """(mu, rho, prior) of one tensor under from_frequentist: uniform draw first (the ctor always makes it), MOPED
    overwrite + one more uniform draw for the prior's own construction when delta is given (linear.py:137-150)."""
    mu, rho = uniform_init(torch.zeros_like(src), torch.zeros_like(src))
    prior = default_mixture_prior()
    if delta is not None:
        mu, rho = src, moped_rho(src, delta)
        uniform_init(torch.zeros_like(src), torch.zeros_like(src))
        prior = gaussian_prior(src.clone(), torch.ones_like(src))
    return mu, rho, prior


def oracle_convert(model: nn.Module, delta: Optional[float], freeze: bool, eps: Optional[EpsSource] = None,
                   all_layers: bool = False):
    """deep-copy + swap every exact-class nn.Linear child for `OracleLinear`
    with MOPED (delta given) or default-uniform init.  `all_layers` also swaps nn.Embedding / nn.LayerNorm (rows
    A9 / A10, composed from the reference's Gaussian arithmetic).
    Ref: bayeformers/__init__.py:50-61; bayeformers/nn/layers/linear.py:106-164."""
    eps = eps or EpsSource()
    new = copy.deepcopy(model)

    def walk(mod):
        for name, child in mod.named_children():
            if child.__class__ is nn.Linear:
                out_f, in_f = child.weight.shape
                has_b = child.bias is not None
                # the ctor always draws the uniform init first (linear.py:137 -> gaussian.py:72-79)
                w_mu, w_rho = uniform_init(torch.zeros(out_f, in_f), torch.zeros(out_f, in_f))
                b_mu = b_rho = None
                if has_b:
                    b_mu, b_rho = uniform_init(torch.zeros(out_f), torch.zeros(out_f))
                w_prior = default_mixture_prior()
                b_prior = default_mixture_prior() if has_b else None
                trainable = True
                if delta is not None:
                    w = child.weight.data
                    w_mu, w_rho = w, moped_rho(w, delta)
                    # the MOPED prior is itself a freshly constructed Gaussian, so it
                    # consumes one more (mu, rho) uniform draw before being overwritten
                    # (linear.py:147-149 -> gaussian.py:72-79)
                    uniform_init(torch.zeros_like(w), torch.zeros_like(w))
                    w_prior = gaussian_prior(w.clone(), torch.ones_like(w))
                    if has_b:
                        b = child.bias.data
                        b_mu, b_rho = b, moped_rho(b, delta)
                        uniform_init(torch.zeros_like(b), torch.zeros_like(b))
                        b_prior = gaussian_prior(b.clone(), torch.ones_like(b))
                    trainable = not freeze
                setattr(mod, name, OracleLinear(w_mu, w_rho, b_mu, b_rho, w_prior, b_prior, eps, trainable))
            elif all_layers and child.__class__ is nn.Embedding:
                trainable = not (delta is not None and freeze)
                w_mu, w_rho, w_prior = _moped_pair(child.weight.data, delta)
                setattr(mod, name, OracleEmbedding(w_mu, w_rho, w_prior, eps, child.padding_idx, trainable))
            elif all_layers and child.__class__ is nn.LayerNorm and child.elementwise_affine:
                trainable = not (delta is not None and freeze)
                w_mu, w_rho, w_prior = _moped_pair(child.weight.data, delta)
                b_mu = b_rho = b_prior = None
                if child.bias is not None:
                    b_mu, b_rho, b_prior = _moped_pair(child.bias.data, delta)
                setattr(mod, name, OracleLayerNorm(child.normalized_shape, child.eps, w_mu, w_rho, b_mu, b_rho,
                                                   w_prior, b_prior, eps, trainable))
            else:
                walk(child)

    walk(new)
    return new


def oracle_layers(model: nn.Module) -> List[nn.Module]:
    return [m for m in model.modules() if isinstance(m, (OracleLinear, _OracleVariational))]


def model_log_prior(model: nn.Module):
    """Python-side sum over Bayesian children of the LAST forward's scalars.
    Ref: bayeformers/nn/model.py:70-78."""
    v = 0.0
    for m in oracle_layers(model):
        v = v + m.log_prior
    return v


def model_log_variational_posterior(model: nn.Module):
    """Ref: bayeformers/nn/model.py:81-89."""
    v = 0.0
    for m in oracle_layers(model):
        v = v + m.log_variational_posterior
    return v


def s_loop_step(model: nn.Module, call: Callable[[nn.Module], torch.Tensor], labels: torch.Tensor,
                samples: int, n_batches: int):
    """The reference's training step pattern: S sequential forwards, mean of
    logits, loss = (lvp - lp)/n_batches + CE(mean logits), backward.
    `call(model)` returns the logits of one forward.
    Ref: examples/bert_glue.py:56-73 (sample_bayesian), :231-239."""
    logits, lps, lqs = [], [], []
    for _ in range(samples):
        logits.append(call(model))
        lps.append(model_log_prior(model))
        lqs.append(model_log_variational_posterior(model))
    mean_logits = torch.stack(logits).mean(0)
    lp = torch.stack([torch.as_tensor(v) for v in lps]).mean()
    lq = torch.stack([torch.as_tensor(v) for v in lqs]).mean()
    nll = F.cross_entropy(mean_logits.view(-1, mean_logits.shape[-1]), labels.view(-1))
    loss = (lq - lp) / n_batches + nll
    loss.backward()
    return loss.detach(), torch.stack(logits).detach(), lp, lq
