"""CPU: pin oracle/ against the fixtures generated from the unmodified
reference (tests/golden/make_golden.py), and the Philox oracle against the
published Random123 known-answer vectors."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import bayes_oracle as O
from oracle import philox_oracle as P


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_gaussian_kat_exact():
    g = load_golden("gaussian.npz")
    for tag in ("kat", "rnd"):
        mu, rho, eps = T(g[f"{tag}_mu"]), T(g[f"{tag}_rho"]), T(g[f"{tag}_eps"])
        assert torch.equal(O.sigma_of(rho), T(g[f"{tag}_sigma"]))
        w = O.gaussian_sample(mu, rho, eps)
        assert torch.equal(w, T(g[f"{tag}_w"]))
        assert torch.equal(O.gaussian_log_prob(w, mu, rho), T(g[f"{tag}_logq"]))
    # SURVEY 8c hand values
    np.testing.assert_allclose(g["kat_sigma"], [0.6931471825, 0.006715348456, 25.0], rtol=1e-7)
    np.testing.assert_allclose(g["kat_w"], [0.6931471825, 0.9932846427, 10.5], rtol=1e-7)
    assert abs(float(g["kat_logq"]) - (-1.7308204174)) < 1e-6
    # float64 closed form agrees with the fp32 reference to fp32 accuracy
    f64 = O.gaussian_log_prob_f64(g["kat_w"], g["kat_mu"], g["kat_rho"])
    assert abs(f64 - float(g["kat_logq"])) < 5e-6


def test_mixture_kat():
    g = load_golden("mixture.npz")
    w = T(g["kat_w"])
    per = O.mixture_log_prob_elementwise(w)
    assert torch.equal(per, T(g["kat_logp_elem"]))
    expect = [4.3903899, 4.3092217, -1.5006596, -1.6133357, -2.1120856, -6.1120858]
    np.testing.assert_allclose(g["kat_logp_elem"][:6], expect, rtol=2e-7)
    f64 = O.mixture_log_prob_elementwise_f64(g["kat_w"][:6])
    np.testing.assert_allclose(f64, g["kat_logp_elem"][:6], rtol=1e-6)
    assert torch.equal(O.mixture_log_prob(T(g["rnd_w"])), T(g["rnd_logp"]))
    pi, s1, s2 = (float(v) for v in g["custom_params"])
    assert torch.equal(O.mixture_log_prob(T(g["rnd_w"]), pi, s1, s2), T(g["custom_logp"]))
    assert float(g["pi"]) == 0.5 and float(g["sigma1"]) == 1.0
    assert float(g["sigma2"]) == np.float32(np.exp(-6))


def test_moped_bit_exact():
    g = load_golden("moped.npz")
    w = T(g["w"])
    for delta in (0.05, 0.1, 0.01):
        rho = O.moped_rho(w, delta)
        assert torch.equal(rho, T(g[f"rho_w_{delta}"]))
        assert np.array_equal(rho.numpy().view(np.uint32), g[f"rho_w_{delta}"].view(np.uint32))
        assert torch.equal(O.moped_rho(w[:, 0].clone(), delta), T(g[f"rho_b_{delta}"]))
        assert torch.equal(w, T(g[f"mu_w_{delta}"]))
    r = g["rho_w_0.05"].ravel()
    assert r[0] == 0.0 and r[1] == 0.0  # w=0 and 1e-7 -> -inf -> 0
    assert np.isposinf(r[9])  # w=2000: +inf is left alone by the reference


@pytest.mark.parametrize("tag", ["default", "moped", "moped_frozen", "nobias", "moped_tc", "moped_frozen_tc",
                                 "moped_frozen_tiles"])
def test_linear_forward_backward(tag):
    g = load_golden("linear.npz")
    in_f, out_f, batch, delta, freeze, bias = g[f"{tag}_meta"]
    bias = bool(bias)
    w_mu = T(g[f"{tag}_w_mu"]).clone().requires_grad_(not freeze)
    w_rho = T(g[f"{tag}_w_rho"]).clone().requires_grad_()
    b_mu = b_rho = eps_b = None
    if bias:
        b_mu = T(g[f"{tag}_b_mu"]).clone().requires_grad_(not freeze)
        b_rho = T(g[f"{tag}_b_rho"]).clone().requires_grad_()
        eps_b = T(g[f"{tag}_eps_b"])
    if delta > 0:
        # MOPED values themselves
        assert torch.equal(T(g[f"{tag}_w_mu"]), T(g[f"{tag}_w0"]))
        assert torch.equal(O.moped_rho(T(g[f"{tag}_w0"]), float(delta)), T(g[f"{tag}_w_rho"]))
        assert torch.equal(T(g[f"{tag}_prior_mu"]), T(g[f"{tag}_w0"]))
        assert np.all(g[f"{tag}_prior_rho"] == 1.0)
        assert bool(g[f"{tag}_mu_requires_grad"][0]) == (not freeze)
        w_prior = O.gaussian_prior(T(g[f"{tag}_w0"]), torch.ones_like(w_mu))
        b_prior = O.gaussian_prior(T(g[f"{tag}_b0"]), torch.ones_like(b_mu)) if bias else None
    else:
        w_prior = O.default_mixture_prior()
        b_prior = O.default_mixture_prior() if bias else None
    x = T(g[f"{tag}_x"]).clone().requires_grad_()
    y, lp, lq, _, _ = O.linear_forward(x, w_mu, w_rho, b_mu, b_rho, T(g[f"{tag}_eps_w"]), eps_b, w_prior, b_prior)
    assert torch.equal(y, T(g[f"{tag}_y"]))
    assert torch.equal(lp, T(g[f"{tag}_log_prior"]))
    assert torch.equal(lq, T(g[f"{tag}_log_q"]))
    assert not lp.requires_grad and not lq.requires_grad  # quirk Q1
    y.backward(T(g[f"{tag}_gy"]))
    assert torch.equal(x.grad, T(g[f"{tag}_g_x"]))
    assert torch.equal(w_rho.grad, T(g[f"{tag}_g_w_rho"]))
    if not freeze:
        assert torch.equal(w_mu.grad, T(g[f"{tag}_g_w_mu"]))
    else:
        assert f"{tag}_g_w_mu" not in g.files
    if bias:
        assert torch.equal(b_rho.grad, T(g[f"{tag}_g_b_rho"]))
    # closed form of SURVEY A6: drho = dW * eps * sigmoid(rho)
    if not freeze:
        closed = g[f"{tag}_g_w_mu"] * g[f"{tag}_eps_w"] / (1.0 + np.exp(-g[f"{tag}_w_rho"].astype(np.float64)))
        assert rel_err(closed, g[f"{tag}_g_w_rho"]) < 1e-6


@pytest.mark.parametrize("tag", ["mixture", "gaussian"])
def test_kl_grad_closed_form(tag):
    g = load_golden("kl_grad.npz")
    c = float(g["kl_weight"])
    if tag == "mixture":
        prior = O.default_mixture_prior()
        prior_t = prior
    else:
        prior = {"kind": "gaussian", "mu": g["gaussian_prior_mu"], "rho": g["gaussian_prior_rho"]}
        prior_t = O.gaussian_prior(T(g["gaussian_prior_mu"]), T(g["gaussian_prior_rho"]))
    mu = T(g["mu"]).clone().requires_grad_()
    rho = T(g["rho"]).clone().requires_grad_()
    lq, lp, _ = O.elbo_terms_with_grad(mu, rho, T(g["eps"]), prior_t)
    assert torch.equal(lq, T(g[f"{tag}_logq"])) and torch.equal(lp, T(g[f"{tag}_logp"]))
    (c * (lq - lp)).backward()
    assert torch.equal(mu.grad, T(g[f"{tag}_g_mu"]))
    assert torch.equal(rho.grad, T(g[f"{tag}_g_rho"]))
    g_mu, g_rho = O.kl_grads_f64(g["mu"], g["rho"], g["eps"], prior, c, -c)
    # fp32 autograd of the mu-gradient cancels two ~eps/sigma terms: loose there
    assert rel_err(g[f"{tag}_g_rho"], g_rho) < 1e-4
    assert rel_err(g[f"{tag}_g_mu"], g_mu) < 5e-3


def test_philox_kat():
    for ctr, key, expect in P.KAT_PHILOX4X32_10:
        out = P.philox4x32_10(*ctr, *key)
        assert [int(o) for o in out] == list(expect)


def test_philox_normal_moments_and_streams():
    from scipy import stats

    x = P.philox_normal(1 << 18, seed=20260101, step=3, tensor_id=5, sample_id=1)
    assert abs(x.mean()) < 4 / np.sqrt(x.size)
    assert abs(x.var() - 1) < 0.01
    assert stats.kstest(x, "norm").pvalue > 1e-3
    y = P.philox_normal(1 << 18, seed=20260101, step=3, tensor_id=5, sample_id=2)
    assert abs(np.corrcoef(x, y)[0, 1]) < 0.01
    # prefix property: a shorter request is a prefix of a longer one
    assert np.array_equal(P.philox_normal(1001, 7, 0, 0, 0), P.philox_normal(4096, 7, 0, 0, 0)[:1001])


def test_dropout_mask_contract():
    """keep = u16 >= round(p * 65536) on the 16-bit halves of the Philox words (first word pinned by the KAT)."""
    m = P.dropout_keep_mask(1 << 20, 0.1, seed=11, step=2, site_id=5)
    assert abs(m.mean() - 0.9) < 5 * np.sqrt(0.09 / m.size) + 1e-5
    r = P.philox4x32_10(0, 0, 0x80000000 | 5, 2, 11, 0)
    thr = int(np.floor(np.float32(0.1).astype(np.float64) * 65536 + 0.5))
    first = [(int(r[i]) >> sh) & 0xFFFF for i in range(4) for sh in (0, 16)]
    assert list(m[:8]) == [int(u >= thr) for u in first]
    assert np.array_equal(P.dropout_keep_mask(1001, 0.1, 11, 2, 5), m[:1001])  # prefix property
    assert not np.array_equal(P.dropout_keep_mask(4096, 0.1, 11, 3, 5), m[:4096])  # next step: new mask
    assert P.dropout_keep_mask(100, 0.0, 1, 1, 1).all()
    # known answers (the CUDA kernels reproduce the oracle bit for bit on the B200: tests/test_gpu_parity.py); pinned
    # here so the mask contract cannot drift without a test noticing
    kat = {(0.1, 0xDEADBEEFCAFEF00D, 9, 77): "1111111001111111011111111111011111111111011110111111101111101011",
           (0.5, 1234, 0, 1): "1111010010110101111111001100111100100001110011001000011110101011"}
    for (p_, seed, step, site), bits in kat.items():
        assert "".join(map(str, P.dropout_keep_mask(64, p_, seed, step, site))) == bits


@pytest.mark.skipif(not os.path.isdir("/root/reference/bayeformers"), reason="live reference only in the build container")
def test_oracle_convert_matches_live_reference():
    import sys

    sys.path.insert(0, "/root/reference")
    try:
        from bayeformers import to_bayesian as ref_to_bayesian
        import bayeformers.nn as rbnn
    finally:
        sys.path.remove("/root/reference")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(10, 12), torch.nn.Tanh(), torch.nn.Linear(12, 3, bias=False))
    torch.manual_seed(5)
    ref = ref_to_bayesian(net, delta=0.05, freeze=True)
    rng_ref = torch.get_rng_state()
    torch.manual_seed(5)
    ora = O.oracle_convert(net, delta=0.05, freeze=True)
    assert torch.equal(rng_ref, torch.get_rng_state())  # same RNG consumption (quirk Q12)
    rl = [m for m in ref.modules() if isinstance(m, rbnn.Linear)]
    ol = O.oracle_layers(ora)
    assert len(rl) == len(ol) == 2
    for r, o in zip(rl, ol):
        assert torch.equal(r.weight.mu, o.w_mu) and torch.equal(r.weight.rho, o.w_rho)
        assert r.weight.mu.requires_grad == o.w_mu.requires_grad
    # same eps -> same outputs
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(5, 10, generator=gen)
    eps = []
    for r in rl:
        eps.append(torch.randn(r.weight.mu.shape, generator=gen))
        if isinstance(r.bias, rbnn.Gaussian):
            eps.append(torch.randn(r.bias.mu.shape, generator=gen))

    class Fixed:
        def __init__(self, e):
            self.e = e

        def sample(self, size):
            return self.e

    it = iter(eps)
    for r in rl:
        r.weight.normal = Fixed(next(it))
        if isinstance(r.bias, rbnn.Gaussian):
            r.bias.normal = Fixed(next(it))
    src = O.EpsSource(preset=list(eps))
    for o in ol:
        o.eps = src
    assert torch.equal(ref(x), ora(x))
    assert torch.equal(ref.log_prior(), O.model_log_prior(ora))
    assert torch.equal(ref.log_variational_posterior(), O.model_log_variational_posterior(ora))


def test_attention_keep_mask_oracle_rate_and_determinism():
    """oracle.attention_keep_mask (the contract the attention kernels are checked against on the GPU): keep rate of the
    quantised threshold, determinism, independence of consecutive steps, everything kept at p = 0."""
    from oracle import philox_oracle as P
    m = P.attention_keep_mask(4, 3, 128, 0.1, 11, 5, 3)
    thr = round(0.1 * 65536)
    n = m.size
    assert m.shape == (4, 3, 128, 128) and m.dtype == np.uint8
    assert abs(m.mean() - (1 - thr / 65536)) < 5 * np.sqrt(0.09 / n)
    assert np.array_equal(m, P.attention_keep_mask(4, 3, 128, 0.1, 11, 5, 3))
    m2 = P.attention_keep_mask(4, 3, 128, 0.1, 11, 6, 3).astype(np.float64)
    assert abs(np.corrcoef(m.reshape(-1).astype(np.float64), m2.reshape(-1))[0, 1]) < 5 / np.sqrt(n)
    assert P.attention_keep_mask(1, 1, 16, 0.0, 1, 1, 1).all()
