"""GPU parity tests proper (run with `-m gpu` on the B200 box): the CUDA path,
called through the C-ABI, against the oracle and the committed golden vectors
generated from the unmodified reference.

Tolerances are the north star's: injected-eps logits and log-prob sums within
1e-5 relative in fp32 mode, 1e-2 in bf16 GEMM mode; MOPED mu/rho bit-exact;
Philox eps moments by statistical test.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import bayes_oracle as O
from oracle import philox_oracle as P

import bayeformers_b200 as bf
import bayeformers_b200.nn as bnn
from bayeformers_b200 import _lib, ops
from bayeformers_b200._lib import BF_BF16, BF_F32, BF_PRIOR_GAUSSIAN, BF_PRIOR_MIXTURE, BF_PRIOR_NONE

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP32_TOL = 1e-5
BF16_TOL = 1e-2


def T(a, dev=DEV):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


class FixedEps:
    """Same stub the golden generator puts on the reference's Gaussian.normal."""

    def __init__(self, queue):
        self.queue = [torch.as_tensor(q) for q in queue]

    def sample(self, size):
        e = self.queue.pop(0)
        assert tuple(e.shape) == tuple(size)
        return e


def run_sample_kl(mu, rho, prior, S, eps=None, w_dtype=torch.float32, seed=0, step=0, tid=0):
    logq = torch.empty(S, device=DEV)
    logp = torch.empty(S, device=DEV)
    w = ops.sample_kl_forward(mu, rho, prior, ops.StreamSpec(seed, tid, step, eps), S, w_dtype, logq, logp, False)
    return w, logq, logp


# ------------------------------------------------------------------ Philox
def test_philox_matches_oracle_contract():
    for (n, seed, step, tid, sid) in [(4096, 0, 0, 0, 0), (1001, 0xDEADBEEFCAFEF00D, 7, 123, 3), (5, 1, 2, 3, 4)]:
        got = ops.philox_normal(n, seed, step, tid, sid, DEV).cpu().numpy()
        want = P.philox_normal(n, seed, step, tid, sid)
        # integer Philox stage must be bit-exact (any bit error scrambles every value);
        # Box-Muller runs on MUFU approximations: absolute error ~1e-6
        assert np.max(np.abs(got - want)) < 2e-5


def test_philox_moments_ks_and_independence():
    from scipy import stats
    n = 1 << 22
    x = ops.philox_normal(n, 20260101, 0, 11, 0, DEV).double().cpu().numpy()
    assert abs(x.mean()) < 4.5 / np.sqrt(n)
    assert abs(x.var() - 1.0) < 4.5 * np.sqrt(2.0 / n)
    assert abs(stats.skew(x)) < 4.5 * np.sqrt(6.0 / n)
    assert abs(stats.kurtosis(x)) < 4.5 * np.sqrt(24.0 / n)
    assert stats.kstest(x[: 1 << 20], "norm").pvalue > 1e-3
    # streams differing in exactly one coordinate are uncorrelated
    for other in [dict(sample_id=1), dict(tensor_id=12), dict(step=1), dict(seed=20260102)]:
        kw = dict(seed=20260101, step=0, tensor_id=11, sample_id=0)
        kw.update(other)
        y = ops.philox_normal(n, kw["seed"], kw["step"], kw["tensor_id"], kw["sample_id"], DEV).double().cpu().numpy()
        assert abs(np.corrcoef(x, y)[0, 1]) < 5.0 / np.sqrt(n)
    # lag-1 autocorrelation inside a stream (pairs come from one Box-Muller)
    assert abs(np.corrcoef(x[:-1], x[1:])[0, 1]) < 5.0 / np.sqrt(n)
    # replicas: same coordinates -> bit-identical stream
    z = ops.philox_normal(n, 20260101, 0, 11, 0, DEV).double().cpu().numpy()
    assert np.array_equal(x, z)


# ------------------------------------------------------------------ sample + KL forward
def test_gaussian_kat_and_random_golden():
    g = load_golden("gaussian.npz")
    for tag in ("kat", "rnd"):
        mu, rho, eps = T(g[f"{tag}_mu"]), T(g[f"{tag}_rho"]), T(g[f"{tag}_eps"])
        w, logq, _ = run_sample_kl(mu, rho, ops.PriorSpec(BF_PRIOR_NONE), 1, eps[None])
        assert rel_err(w[0].cpu().numpy(), g[f"{tag}_w"]) < 1e-6
        assert abs(float(logq[0]) - float(g[f"{tag}_logq"])) <= FP32_TOL * abs(float(g[f"{tag}_logq"]))


def test_mixture_prior_golden():
    g = load_golden("mixture.npz")
    for w_np, want, (pi, s1, s2) in [(g["rnd_w"], g["rnd_logp"], (0.5, 1.0, float(np.float32(np.exp(-6))))),
                                     (g["rnd_w"], g["custom_logp"], tuple(float(v) for v in g["custom_params"]))]:
        mu = T(w_np)
        rho = torch.full_like(mu, -40.0)  # sigma ~ 4e-18: w == mu exactly
        _, _, logp = run_sample_kl(mu, rho, ops.PriorSpec(BF_PRIOR_MIXTURE, pi, s1, s2), 1, torch.zeros_like(mu)[None])
        assert abs(float(logp[0]) - float(want)) <= FP32_TOL * abs(float(want))
    # per-element KATs incl. the fp32 underflow region (w = 14)
    for wv, want in zip(g["kat_w"], g["kat_logp_elem"]):
        mu = T(np.array([wv], dtype=np.float32))
        _, _, logp = run_sample_kl(mu, torch.full_like(mu, -40.0), ops.PriorSpec(BF_PRIOR_MIXTURE, 0.5, 1.0, float(np.float32(np.exp(-6)))), 1, torch.zeros(1, 1, device=DEV))
        tol = 1e-4 if abs(wv) > 10 else 2e-6
        assert abs(float(logp[0]) - float(want)) <= tol * abs(float(want))


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 4099, 65536 + 7])
@pytest.mark.parametrize("S", [1, 3, 8, 11])
@pytest.mark.parametrize("prior_kind", ["mixture", "gaussian"])
def test_sample_kl_forward_vs_oracle(n, S, prior_kind):
    gen = torch.Generator().manual_seed(n * 31 + S)
    mu = torch.empty(n).uniform_(-0.2, 0.2, generator=gen)
    rho = torch.empty(n).uniform_(-6.0, -2.0, generator=gen)
    eps = torch.randn(S, n, generator=gen)
    if prior_kind == "mixture":
        prior_o = O.default_mixture_prior()
        prior_k = ops.PriorSpec(BF_PRIOR_MIXTURE, 0.5, 1.0, float(np.float32(np.exp(-6))))
    else:
        pm = mu + 0.01 * torch.randn(n, generator=gen)
        pr = torch.ones(n)
        prior_o = O.gaussian_prior(pm, pr)
        prior_k = ops.PriorSpec(BF_PRIOR_GAUSSIAN, mu=pm.to(DEV), rho=pr.to(DEV))
    w, logq, logp = run_sample_kl(mu.to(DEV), rho.to(DEV), prior_k, S, eps.to(DEV))
    wb, logq_b, logp_b = run_sample_kl(mu.to(DEV), rho.to(DEV), prior_k, S, eps.to(DEV), w_dtype=torch.bfloat16)
    for s in range(S):
        w_o = O.gaussian_sample(mu, rho, eps[s])
        assert rel_err(w[s].cpu().numpy(), w_o.numpy()) < 1e-6
        assert rel_err(wb[s].float().cpu().numpy(), w_o.numpy()) < 4e-3  # bf16 rounding of the stored sample only
        lq, lp = float(O.gaussian_log_prob(w_o, mu, rho)), float(O.prior_log_prob(w_o, prior_o))
        assert abs(float(logq[s]) - lq) <= FP32_TOL * max(abs(lq), 1e-3)
        assert abs(float(logp[s]) - lp) <= FP32_TOL * max(abs(lp), 1e-3)
    # log-probs are computed from the fp32 sample even when w is stored as bf16 (a bf16-rounded sample would move them
    # by ~4e-3 relative).  Not bitwise: for tiny n the two dtypes can take different kernels (16-byte alignment of the
    # sample rows), whose fp32 partial sums are grouped differently.
    assert torch.allclose(logq, logq_b, rtol=2e-6, atol=0) and torch.allclose(logp, logp_b, rtol=2e-6, atol=0)


def test_sample_kl_edge_cases():
    # empty tensor: sums are exactly 0
    z = torch.empty(0, device=DEV)
    _, lq, lp = run_sample_kl(z, z, ops.PriorSpec(BF_PRIOR_MIXTURE, 0.5, 1.0, 0.0025), 2)
    assert torch.all(lq == 0) and torch.all(lp == 0)
    # accumulate (weight then bias, linear.py:99-102) and w_out == NULL
    mu = torch.randn(100, device=DEV) * 0.1
    rho = torch.full((100,), -4.0, device=DEV)
    pr = ops.PriorSpec(BF_PRIOR_MIXTURE, 0.5, 1.0, 0.0025)
    _, lq1, lp1 = run_sample_kl(mu, rho, pr, 2, seed=5)
    logq, logp = lq1.clone(), lp1.clone()
    ops.sample_kl_forward(mu, rho, pr, ops.StreamSpec(5, 0, 0), 2, torch.float32, logq, logp, True, want_w=False)
    assert torch.allclose(logq, 2 * lq1) and torch.allclose(logp, 2 * lp1)
    # misaligned views take the scalar path and agree with the vector path
    big_mu, big_rho = torch.randn(1001, device=DEV), torch.full((1001,), -3.0, device=DEV)
    w_a, lq_a, _ = run_sample_kl(big_mu[1:].clone(), big_rho[1:].clone(), pr, 1, seed=9)
    w_b, lq_b, _ = run_sample_kl(big_mu[1:], big_rho[1:], pr, 1, seed=9)
    assert torch.equal(w_a, w_b) and torch.allclose(lq_a, lq_b, rtol=1e-6)


def test_sample_kl_deterministic_and_chunk_invariant():
    n, S = 1 << 20, 11
    mu = torch.randn(n, device=DEV) * 0.05
    rho = torch.empty(n, device=DEV).uniform_(-6, -3)
    pr = ops.PriorSpec(BF_PRIOR_MIXTURE, 0.5, 1.0, float(np.float32(np.exp(-6))))
    w1, q1, p1 = run_sample_kl(mu, rho, pr, S, seed=42, step=3, tid=9)
    w2, q2, p2 = run_sample_kl(mu, rho, pr, S, seed=42, step=3, tid=9)
    assert torch.equal(w1, w2) and torch.equal(q1, q2) and torch.equal(p1, p2)  # run-to-run bitwise
    # sample s of an S-sample launch == the same sample id drawn alone by the eps stream
    eps5 = ops.philox_normal(n, 42, 3, 9, 5, DEV)
    w5 = O.gaussian_sample(mu.cpu(), rho.cpu(), eps5.cpu())
    assert rel_err(w1[5].cpu().numpy(), w5.numpy()) < 1e-6


# ------------------------------------------------------------------ backward
@pytest.mark.parametrize("tag", ["mixture", "gaussian"])
def test_kl_grad_against_reference_autograd_and_f64(tag):
    g = load_golden("kl_grad.npz")
    c = float(g["kl_weight"])
    mu, rho, eps = T(g["mu"]), T(g["rho"]), T(g["eps"])
    if tag == "mixture":
        pk = ops.PriorSpec(BF_PRIOR_MIXTURE, 0.5, 1.0, float(np.float32(np.exp(-6))))
        prior64 = O.default_mixture_prior()
    else:
        pk = ops.PriorSpec(BF_PRIOR_GAUSSIAN, mu=T(g["gaussian_prior_mu"]), rho=T(g["gaussian_prior_rho"]))
        prior64 = {"kind": "gaussian", "mu": g["gaussian_prior_mu"], "rho": g["gaussian_prior_rho"]}
    glq = torch.tensor([c], device=DEV)
    glp = torch.tensor([-c], device=DEV)
    g_mu, g_rho = ops.sample_kl_backward(None, mu, rho, pk, ops.StreamSpec(eps=eps[None]), 1, glq, glp, True)
    want_mu, want_rho = O.kl_grads_f64(g["mu"], g["rho"], g["eps"], prior64, c, -c)
    assert rel_err(g_rho.cpu().numpy(), want_rho) < 1e-5
    assert rel_err(g_mu.cpu().numpy(), want_mu) < 1e-5
    # and against the reference's own fp32 autograd (rho-gradient; its mu-gradient cancels badly in fp32)
    assert rel_err(g_rho.cpu().numpy(), g[f"{tag}_g_rho"]) < 1e-4


def test_backward_recomputes_the_same_eps():
    n, S = 10007, 5
    mu = (torch.randn(n) * 0.1).to(DEV)
    rho = torch.empty(n).uniform_(-5, -2).to(DEV)
    gw = torch.randn(S, n, device=DEV)
    st = ops.StreamSpec(seed=314159, tensor_id=21, step=6)
    g_mu, g_rho = ops.sample_kl_backward(gw, mu, rho, ops.PriorSpec(), st, S, None, None, True)
    eps = torch.stack([ops.philox_normal(n, 314159, 6, 21, s, DEV) for s in range(S)])
    want_rho = (gw * eps).sum(0).double() * torch.sigmoid(rho.double())
    assert rel_err(g_rho.cpu().numpy(), want_rho.cpu().numpy()) < 1e-6
    assert rel_err(g_mu.cpu().numpy(), gw.sum(0).cpu().numpy()) < 1e-6
    gwb = gw.to(torch.bfloat16)
    _, g_rho_b = ops.sample_kl_backward(gwb, mu, rho, ops.PriorSpec(), st, S, None, None, False)
    want_b = (gwb.float() * eps).sum(0).double() * torch.sigmoid(rho.double())
    assert rel_err(g_rho_b.cpu().numpy(), want_b.cpu().numpy()) < 1e-6


# ------------------------------------------------------------------ Linear layer vs golden (reference outputs)
def _layer_from_golden(g, tag, gemm_dtype):
    in_f, out_f, batch, delta, freeze, bias = g[f"{tag}_meta"]
    bias, freeze = bool(bias), bool(freeze)
    lin = torch.nn.Linear(int(in_f), int(out_f), bias=bias)
    lin.weight.data = torch.from_numpy(g[f"{tag}_w0"]).clone()
    if bias:
        lin.bias.data = torch.from_numpy(g[f"{tag}_b0"]).clone()
    layer = bnn.Linear.from_frequentist(lin, delta=(float(delta) if delta > 0 else None), freeze=freeze)
    if delta <= 0:  # the reference kept its own uniform draw: load it
        layer.weight.mu.data = torch.from_numpy(g[f"{tag}_w_mu"]).clone()
        layer.weight.rho.data = torch.from_numpy(g[f"{tag}_w_rho"]).clone()
        if bias:
            layer.bias.mu.data = torch.from_numpy(g[f"{tag}_b_mu"]).clone()
            layer.bias.rho.data = torch.from_numpy(g[f"{tag}_b_rho"]).clone()
    else:  # MOPED must reproduce the reference bit for bit
        assert torch.equal(layer.weight.rho.data, torch.from_numpy(g[f"{tag}_w_rho"]))
    layer = layer.to(DEV)
    layer.gemm_dtype = gemm_dtype
    layer.weight.normal = FixedEps([g[f"{tag}_eps_w"]])
    if bias:
        layer.bias.normal = FixedEps([g[f"{tag}_eps_b"]])
    return layer, bias, freeze


LINEAR_TAGS = ["default", "moped", "moped_frozen", "nobias", "moped_tc", "moped_frozen_tc", "moped_frozen_tiles"]


@pytest.mark.parametrize("tag", LINEAR_TAGS)
def test_linear_fp32_matches_reference(tag):
    g = load_golden("linear.npz")
    layer, bias, freeze = _layer_from_golden(g, tag, torch.float32)
    x = T(g[f"{tag}_x"]).requires_grad_()
    y = layer(x)
    assert rel_err(y.detach().cpu().numpy(), g[f"{tag}_y"]) < FP32_TOL
    assert abs(float(layer.log_prior) - float(g[f"{tag}_log_prior"])) <= FP32_TOL * abs(float(g[f"{tag}_log_prior"]))
    assert abs(float(layer.log_variational_posterior) - float(g[f"{tag}_log_q"])) <= FP32_TOL * abs(float(g[f"{tag}_log_q"]))
    assert layer.log_prior.dim() == 0 and not layer.log_prior.requires_grad  # quirks Q1/Q2
    y.backward(T(g[f"{tag}_gy"]))
    assert rel_err(x.grad.cpu().numpy(), g[f"{tag}_g_x"]) < FP32_TOL
    assert rel_err(layer.weight.rho.grad.cpu().numpy(), g[f"{tag}_g_w_rho"]) < FP32_TOL
    if freeze:
        assert layer.weight.mu.grad is None
    else:
        assert rel_err(layer.weight.mu.grad.cpu().numpy(), g[f"{tag}_g_w_mu"]) < FP32_TOL
    if bias:
        assert rel_err(layer.bias.rho.grad.cpu().numpy(), g[f"{tag}_g_b_rho"]) < FP32_TOL
        if not freeze:
            assert rel_err(layer.bias.mu.grad.cpu().numpy(), g[f"{tag}_g_b_mu"]) < FP32_TOL


@pytest.mark.parametrize("tag", LINEAR_TAGS)
def test_linear_bf16_matches_reference(tag):
    """bf16 GEMM mode against outputs of the unmodified reference.  The `*_tc` / `*_tiles` cases are MOPED layers
    (Gaussian prior, frozen or trainable mu) whose shapes are multiples of 8: forward, dgrad and the fused
    variational wgrad all run on the tcgen05 kernels (asserted through the launch timers' names)."""
    g = load_golden("linear.npz")
    layer, bias, freeze = _layer_from_golden(g, tag, torch.bfloat16)
    N, K = layer.weight.mu.shape
    on_tc = tag in ("default", "nobias", "moped_tc", "moped_frozen_tc", "moped_frozen_tiles")
    if on_tc:  # 8-aligned shapes: these really run the tcgen05 kernels
        assert ops.tc_eligible(N, K)
    x = T(g[f"{tag}_x"]).requires_grad_()
    ops.enable_kernel_timing(True)
    try:
        y = layer(x)
        y.backward(T(g[f"{tag}_gy"]), retain_graph=False)
        torch.cuda.synchronize()
        ran = set(ops.kernel_timing_summary())
    finally:
        ops.enable_kernel_timing(False)
    if on_tc:
        assert {"gemm_fwd_tc", "gemm_dgrad_tc", "gemm_wgrad_fused_tc"} <= ran, ran
    x.grad = None
    layer.weight.rho.grad = layer.weight.mu.grad = None
    if bias:
        layer.bias.rho.grad = layer.bias.mu.grad = None
    layer.weight.normal = FixedEps([g[f"{tag}_eps_w"]])
    if bias:
        layer.bias.normal = FixedEps([g[f"{tag}_eps_b"]])
    y = layer(x)
    assert y.dtype == torch.float32
    assert rel_err(y.detach().cpu().numpy(), g[f"{tag}_y"]) < BF16_TOL
    # the log-probs do not depend on the GEMM dtype
    assert abs(float(layer.log_prior) - float(g[f"{tag}_log_prior"])) <= FP32_TOL * abs(float(g[f"{tag}_log_prior"]))
    y.backward(T(g[f"{tag}_gy"]))
    assert rel_err(x.grad.cpu().numpy(), g[f"{tag}_g_x"]) < BF16_TOL
    assert rel_err(layer.weight.rho.grad.cpu().numpy(), g[f"{tag}_g_w_rho"]) < BF16_TOL
    if not freeze:
        assert rel_err(layer.weight.mu.grad.cpu().numpy(), g[f"{tag}_g_w_mu"]) < BF16_TOL
    if bias:
        assert rel_err(layer.bias.rho.grad.cpu().numpy(), g[f"{tag}_g_b_rho"]) < BF16_TOL


class TinyMLP(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.body = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.ReLU(), torch.nn.Linear(16, 8, bias=False))
        self.head = torch.nn.Linear(8, 3)

    def forward(self, x):
        return self.head(torch.relu(self.body(x)))


@pytest.mark.parametrize("tag,kw", [("plain", {}), ("moped", {"delta": 0.05, "freeze": True}),
                                    ("moped_unfrozen", {"delta": 0.1})])
def test_model_forward_matches_reference(tag, kw):
    g = load_golden("to_bayesian.npz")
    torch.manual_seed(21)
    m = TinyMLP()
    torch.manual_seed(22)
    bm = bf.to_bayesian(m, **kw).to(DEV)
    i = 0
    for mod in bm.bayesian_children:
        mod.weight.normal = FixedEps([g[f"{tag}_eps{i}"]]); i += 1
        if isinstance(mod.bias, bnn.Gaussian):
            mod.bias.normal = FixedEps([g[f"{tag}_eps{i}"]]); i += 1
    y = bm(T(g[f"{tag}_x"]))
    assert rel_err(y.detach().cpu().numpy(), g[f"{tag}_y"]) < FP32_TOL
    lp, lq = bm.log_prior(), bm.log_variational_posterior()
    assert lp.dim() == 0 and not lp.requires_grad
    assert abs(float(lp) - float(g[f"{tag}_log_prior"])) <= FP32_TOL * abs(float(g[f"{tag}_log_prior"]))
    assert abs(float(lq) - float(g[f"{tag}_log_q"])) <= FP32_TOL * abs(float(g[f"{tag}_log_q"]))


# ------------------------------------------------------------------ tiny BERT through the reference S-loop
def _tiny_bert(g, gemm_dtype):
    from transformers import BertConfig, BertForSequenceClassification
    v, h, L, heads, ff, pos, nl = (int(a) for a in g["cfg"])
    cfg = BertConfig(vocab_size=v, hidden_size=h, num_hidden_layers=L, num_attention_heads=heads,
                     intermediate_size=ff, max_position_embeddings=pos, num_labels=nl)
    model = BertForSequenceClassification(cfg).eval()
    sd = {k[len("freq::"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("freq::")}
    model.load_state_dict(sd, strict=True)
    bm = bf.to_bayesian(model, delta=0.05, freeze=True, gemm_dtype=gemm_dtype).eval().to(DEV)
    layers = [m for m in bm.modules() if isinstance(m, bnn.Linear)]
    names = [n for n, m in bm.named_modules() if isinstance(m, bnn.Linear)]
    assert names == [str(s) for s in g["layer_names"]]
    return bm, layers


@pytest.mark.parametrize("mode", ["loop", "folded"])
def test_tiny_bert_s_loop_and_folded_match_reference(mode):
    g = load_golden("tiny_bert.npz")
    S, n_batches = int(g["S"]), int(g["n_batches"])
    bm, layers = _tiny_bert(g, "fp32")
    ids, labels = T(g["ids"]), T(g["labels"])
    if mode == "loop":
        for i, l in enumerate(layers):
            l.weight.normal = FixedEps(list(g[f"eps_w{i}"]))
            l.bias.normal = FixedEps(list(g[f"eps_b{i}"]))
        logits, lps, lqs = [], [], []
        for s in range(S):  # the reference's own usage pattern, unchanged (README.md:62-65)
            logits.append(bm(input_ids=ids).logits)
            lps.append(bm.log_prior())
            lqs.append(bm.log_variational_posterior())
        raw, lp_s, lq_s = torch.stack(logits), torch.stack(lps), torch.stack(lqs)
    else:
        for i, l in enumerate(layers):
            l.weight.normal = FixedEps(list(g[f"eps_w{i}"]))
            l.bias.normal = FixedEps(list(g[f"eps_b{i}"]))
        with bf.mc_samples(S):
            out = bm(input_ids=ids.repeat(S, 1)).logits
        raw = out.view(S, ids.shape[0], -1)
        lp_s, lq_s = bm.log_prior(), bm.log_variational_posterior()
        assert lp_s.shape == (S,)
    assert rel_err(raw.detach().cpu().numpy(), g["logits"]) < 5 * FP32_TOL  # 2 encoder layers deep
    assert rel_err(lp_s.cpu().numpy(), g["log_prior"]) < FP32_TOL
    assert rel_err(lq_s.cpu().numpy(), g["log_q"]) < FP32_TOL
    nll = torch.nn.functional.cross_entropy(raw.mean(0), labels)
    loss = (lq_s.mean() - lp_s.mean()) / n_batches + nll
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    loss.backward()
    scale = max(float(np.abs(g[f"g_b_rho{i}"]).max()) for i in range(len(layers)))
    for i, l in enumerate(layers):
        assert rel_err(l.weight.rho.grad.cpu().numpy(), g[f"g_w_rho{i}"]) < 2e-4, i
        want_b = g[f"g_b_rho{i}"]
        if float(np.abs(want_b).max()) < 1e-9 * scale:
            # key-projection biases: softmax is shift invariant, the true gradient is 0 and
            # both sides hold rounding noise (~1e-17)
            assert float(l.bias.rho.grad.abs().max()) < 1e-7 * scale, i
        else:
            assert rel_err(l.bias.rho.grad.cpu().numpy(), want_b) < 2e-4, i
        assert l.weight.mu.grad is None


def test_tiny_bert_bf16_mode_within_1e2():
    g = load_golden("tiny_bert.npz")
    S = int(g["S"])
    bm, layers = _tiny_bert(g, "bf16")
    for i, l in enumerate(layers):
        l.weight.normal = FixedEps(list(g[f"eps_w{i}"]))
        l.bias.normal = FixedEps(list(g[f"eps_b{i}"]))
    with bf.mc_samples(S):
        out = bm(input_ids=T(g["ids"]).repeat(S, 1)).logits
    raw = out.view(S, -1, out.shape[-1])
    assert rel_err(raw.detach().float().cpu().numpy(), g["logits"]) < BF16_TOL
    assert rel_err(bm.log_prior().cpu().numpy(), g["log_prior"]) < FP32_TOL



def test_tiny_bert_qa_span_head_matches_oracle_loop():
    """BASELINE configs[3] shape in miniature: BertForQuestionAnswering (span head, start/end logits),
    S samples folded through harness.sample_bayesian vs the oracle's sequential S-loop on the CPU
    (pattern of examples/bert_squad.py:190-212) with the same eps."""
    from transformers import BertConfig, BertForQuestionAnswering
    torch.manual_seed(7)
    cfg = BertConfig(vocab_size=120, hidden_size=64, num_hidden_layers=2, num_attention_heads=4,
                     intermediate_size=128, max_position_embeddings=64)
    model = BertForQuestionAnswering(cfg).eval()
    gen = torch.Generator().manual_seed(8)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("bias"):
                p.add_(torch.randn(p.shape, generator=gen) * 0.02)
    S, B, Tn = 3, 2, 24
    ids = torch.randint(0, 120, (B, Tn), generator=gen)
    lin_names = [n for n, m in model.named_modules() if m.__class__ is torch.nn.Linear]
    shapes = {n: dict(model.named_modules())[n] for n in lin_names}
    eps = {(s, n): (torch.randn(shapes[n].weight.shape, generator=gen), torch.randn(shapes[n].bias.shape, generator=gen))
           for s in range(S) for n in lin_names}
    # ours: folded forward
    bm = bf.to_bayesian(model, delta=0.05, freeze=True).eval().to(DEV)
    for n, m in bm.model.named_modules():
        if isinstance(m, bnn.Linear):
            m.weight.normal = FixedEps([eps[(s, n)][0] for s in range(S)])
            m.bias.normal = FixedEps([eps[(s, n)][1] for s in range(S)])
    (raw_start, raw_end), (start, end), lp, lq = bf.sample_bayesian(
        bm, {"input_ids": ids.to(DEV)}, S, select=("start_logits", "end_logits"))
    assert raw_start.shape == (S, B, Tn) and start.shape == (B, Tn)
    # oracle: sequential S-loop, one eps queue in layer execution order (weight then bias)
    queue = [e for s in range(S) for n in lin_names for e in eps[(s, n)]]
    om = O.oracle_convert(model, 0.05, True, eps=O.EpsSource(preset=queue)).eval()
    st, en, lps, lqs = [], [], [], []
    with torch.no_grad():
        for s in range(S):
            out = om(input_ids=ids)
            st.append(out.start_logits), en.append(out.end_logits)
            lps.append(O.model_log_prior(om)), lqs.append(O.model_log_variational_posterior(om))
    assert rel_err(raw_start.detach().cpu().numpy(), torch.stack(st).numpy()) < 5 * FP32_TOL
    assert rel_err(raw_end.detach().cpu().numpy(), torch.stack(en).numpy()) < 5 * FP32_TOL
    assert abs(float(lp) - float(torch.stack(lps).mean())) <= FP32_TOL * abs(float(torch.stack(lps).mean()))
    assert abs(float(lq) - float(torch.stack(lqs).mean())) <= FP32_TOL * abs(float(torch.stack(lqs).mean()))

# ------------------------------------------------------------------ kl_grad extension
def test_kl_grad_true_flows_through_model_scalars():
    torch.manual_seed(0)
    lin = torch.nn.Linear(24, 16)
    layer = bnn.Linear.from_frequentist(lin, delta=0.05).to(DEV)
    layer.kl_grad = True
    gen = torch.Generator().manual_seed(1)
    eps_w, eps_b = torch.randn(16, 24, generator=gen), torch.randn(16, generator=gen)
    layer.weight.normal, layer.bias.normal = FixedEps([eps_w]), FixedEps([eps_b])
    model = bnn.Model(layer)
    x = torch.randn(5, 24, generator=gen).to(DEV)
    y = model(x)
    c = 1.0 / 50
    loss = c * (model.log_variational_posterior() - model.log_prior()) + y.square().sum()
    loss.backward()
    # oracle: the data term through the restated layer + the f64 closed-form KL gradient
    w_mu = lin.weight.detach().clone().requires_grad_()
    w_rho = O.moped_rho(lin.weight.detach(), 0.05).requires_grad_()
    b_mu = lin.bias.detach().clone().requires_grad_()
    b_rho = O.moped_rho(lin.bias.detach(), 0.05).requires_grad_()
    wp = O.gaussian_prior(lin.weight.detach(), torch.ones(16, 24))
    bp = O.gaussian_prior(lin.bias.detach(), torch.ones(16))
    yo, _, _, _, _ = O.linear_forward(x.cpu(), w_mu, w_rho, b_mu, b_rho, eps_w, eps_b, wp, bp)
    yo.square().sum().backward()
    kmu, krho = O.kl_grads_f64(w_mu.detach().numpy(), w_rho.detach().numpy(), eps_w.numpy(),
                               {"kind": "gaussian", "mu": lin.weight.detach().numpy(), "rho": np.ones((16, 24))}, c, -c)
    assert rel_err(layer.weight.rho.grad.cpu().numpy(), w_rho.grad.numpy() + krho) < 1e-5
    assert rel_err(layer.weight.mu.grad.cpu().numpy(), w_mu.grad.numpy() + kmu) < 1e-5
    kmu_b, krho_b = O.kl_grads_f64(b_mu.detach().numpy(), b_rho.detach().numpy(), eps_b.numpy(),
                                   {"kind": "gaussian", "mu": lin.bias.detach().numpy(), "rho": np.ones(16)}, c, -c)
    assert rel_err(layer.bias.rho.grad.cpu().numpy(), b_rho.grad.numpy() + krho_b) < 1e-5


# ------------------------------------------------------------------ tcgen05 contractions vs fp32 math on the same bf16 inputs
@pytest.mark.parametrize("S,M,N,K", [(1, 128, 256, 64), (1, 128, 256, 512), (2, 256, 512, 256), (3, 200, 768, 136),
                                     (1, 1000, 72, 3072), (2, 384, 3072, 768), (4, 1024, 768, 768),
                                     (2, 300, 520, 136), (1, 1000, 776, 1544), (4, 513, 264, 72), (1, 2048, 3072, 768)])
def test_tc_contractions(S, M, N, K):
    lib = _lib.load()
    gen = torch.Generator().manual_seed(S * 1000 + M)
    x = torch.randn(S, M, K, generator=gen).to(DEV).bfloat16()
    w = (torch.randn(S, N, K, generator=gen) * 0.05).to(DEV).bfloat16()
    b = torch.randn(S, N, generator=gen).to(DEV)
    gy = torch.randn(S, M, N, generator=gen).to(DEV).bfloat16()
    st = torch.cuda.current_stream().cuda_stream
    y = torch.empty(S, M, N, device=DEV)
    _lib.check(lib.bf_linear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), S, M, N, K, BF_BF16, BF_F32, st), "fwd")
    want = torch.einsum("smk,snk->smn", x.double(), w.double()) + b.double()[:, None, :]
    assert rel_err(y.cpu().numpy(), want.cpu().numpy()) < 1e-5
    yb = torch.empty(S, M, N, device=DEV, dtype=torch.bfloat16)
    _lib.check(lib.bf_linear_fwd(x.data_ptr(), w.data_ptr(), None, yb.data_ptr(), S, M, N, K, BF_BF16, BF_BF16, st), "fwd bf16")
    assert rel_err(yb.float().cpu().numpy(), (want - b.double()[:, None, :]).cpu().numpy()) < 4e-3
    dx = torch.empty(S, M, K, device=DEV)
    _lib.check(lib.bf_linear_dgrad(gy.data_ptr(), w.data_ptr(), dx.data_ptr(), S, M, N, K, BF_BF16, BF_F32, st), "dgrad")
    want = torch.einsum("smn,snk->smk", gy.double(), w.double())
    assert rel_err(dx.cpu().numpy(), want.cpu().numpy()) < 1e-5
    dw = torch.empty(S, N, K, device=DEV)
    _lib.check(lib.bf_linear_wgrad(gy.data_ptr(), x.data_ptr(), dw.data_ptr(), S, M, N, K, BF_BF16, st), "wgrad")
    want = torch.einsum("smn,smk->snk", gy.double(), x.double())
    assert rel_err(dw.cpu().numpy(), want.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("S,M,N,K,kl", [(1, 256, 128, 256, False), (4, 512, 768, 768, False), (3, 200, 136, 264, True),
                                        (8, 4096, 256, 256, False), (2, 300, 3072, 768, True)])
def test_fused_wgrad_epilogue_equals_unfused_path(S, M, N, K, kl):
    """wgrad with the variational epilogue (eps regenerated from Philox inside the
    GEMM) == plain wgrad followed by the stand-alone backward kernel."""
    lib = _lib.load()
    gen = torch.Generator().manual_seed(M + N)
    x = torch.randn(S, M, K, generator=gen).to(DEV).bfloat16()
    gy = torch.randn(S, M, N, generator=gen).to(DEV).bfloat16()
    mu = (torch.randn(N, K, generator=gen) * 0.05).to(DEV)
    rho = torch.empty(N, K).uniform_(-6, -3, generator=gen).to(DEV)
    prior = ops.PriorSpec(BF_PRIOR_GAUSSIAN, mu=(mu + 0.01).contiguous(), rho=torch.ones_like(mu))
    glq = torch.rand(S, generator=gen).to(DEV) * 0.01 if kl else None
    glp = -torch.rand(S, generator=gen).to(DEV) * 0.01 if kl else None
    stream = ops.StreamSpec(seed=99, tensor_id=5, step=2)
    st = torch.cuda.current_stream().cuda_stream
    dw = torch.empty(S, N, K, device=DEV)
    _lib.check(lib.bf_linear_wgrad(gy.data_ptr(), x.data_ptr(), dw.data_ptr(), S, M, N, K, BF_BF16, st), "wgrad")
    want_mu, want_rho = ops.sample_kl_backward(dw.view(S, -1), mu, rho, prior, stream, S, glq, glp, True)
    g_mu, g_rho = torch.empty_like(mu), torch.empty_like(mu)
    ws = torch.empty(lib.bf_linear_wgrad_fused_workspace_bytes(S, M, N, K, 1), dtype=torch.uint8, device=DEV)

    def fused(acc):
        _lib.check(lib.bf_linear_wgrad_fused(
            gy.data_ptr(), x.data_ptr(), S, M, N, K, BF_BF16, mu.data_ptr(), rho.data_ptr(), prior.kind,
            prior.mu.data_ptr(), prior.rho.data_ptr(), 0.5, 1.0, 1.0, None if glq is None else glq.data_ptr(),
            None if glp is None else glp.data_ptr(), 99, 2, 5, None, g_mu.data_ptr(), g_rho.data_ptr(), acc,
            ws.data_ptr(), st), "wgrad_fused")

    fused(0)
    # both sides are fp32 sums of the same products in different association orders
    # (the fused kernel may cut the reduction into slices): 1e-5 is fp32 noise here
    assert rel_err(g_rho.cpu().numpy(), want_rho.cpu().numpy()) < 1e-5
    assert rel_err(g_mu.cpu().numpy(), want_mu.cpu().numpy()) < 1e-5
    first = g_rho.clone()
    fused(0)
    assert torch.equal(first, g_rho), "fixed-order reduction of the partials must be run-to-run deterministic"
    fused(1)  # accumulate into existing gradients
    assert rel_err(g_rho.cpu().numpy(), 2 * want_rho.cpu().numpy()) < 1e-5








# ------------------------------------------------------------------ bias + GELU epilogue (extension of bnn.Linear)
@pytest.mark.parametrize("S,M,N,K", [(2, 2560, 3072, 768), (4, 1280, 520, 264), (1, 300, 64, 32)])
def test_linear_gelu_fused_vs_separate_and_f64(S, M, N, K):
    """activation="gelu": fused tensor-core epilogue + fused GELU'/bias-grad pass (large shapes) or the composed
    fallback (small ones) against y = gelu(x w^T + b) evaluated in float64, forward and backward."""
    torch.manual_seed(S * 7 + N)
    lin = torch.nn.Linear(K, N)
    with torch.no_grad():
        lin.weight.mul_(2.0)
    layer = bnn.Linear.from_frequentist(lin, delta=0.05).to(DEV)
    layer.gemm_dtype, layer.activation = torch.bfloat16, "gelu"
    gen = torch.Generator().manual_seed(1)
    eps_w = [torch.randn(N, K, generator=gen) for _ in range(S)]
    eps_b = [torch.randn(N, generator=gen) for _ in range(S)]
    layer.weight.normal, layer.bias.normal = FixedEps(eps_w), FixedEps(eps_b)
    x = torch.randn(S * M, K, generator=gen).to(DEV).bfloat16().requires_grad_()
    gy = torch.randn(S * M, N, generator=gen).to(DEV).bfloat16()
    with bf.mc_samples(S):
        y = layer(x)
    assert y.dtype == torch.bfloat16
    y.backward(gy)
    # float64 reference on the device, one sample at a time, from the same bf16-rounded operands
    mu_w, rho_w = layer.weight.mu.detach().double(), layer.weight.rho.detach().double().requires_grad_()
    mu_b, rho_b = layer.bias.mu.detach().double(), layer.bias.rho.detach().double().requires_grad_()
    xr = x.detach().double().requires_grad_()
    outs = []
    for s in range(S):
        w = (mu_w + torch.nn.functional.softplus(rho_w) * eps_w[s].to(DEV).double())
        w = w + (w.detach().float().bfloat16().double() - w.detach())  # bf16-rounded sample, straight-through
        b = mu_b + torch.nn.functional.softplus(rho_b) * eps_b[s].to(DEV).double()
        outs.append(torch.nn.functional.gelu(xr[s * M:(s + 1) * M] @ w.T + b))
    ref = torch.cat(outs)
    ref.backward(gy.double())
    assert rel_err(y.detach().float().cpu().numpy(), ref.detach().cpu().numpy()) < BF16_TOL
    assert rel_err(x.grad.float().cpu().numpy(), xr.grad.cpu().numpy()) < BF16_TOL
    assert rel_err(layer.weight.rho.grad.cpu().numpy(), rho_w.grad.cpu().numpy()) < BF16_TOL
    assert rel_err(layer.bias.rho.grad.cpu().numpy(), rho_b.grad.cpu().numpy()) < BF16_TOL


def test_accelerate_host_fuses_hf_gelu_block():
    """accelerate_host_(fuse_gelu=True) on a HF BERT layer: same logits (bf16 tolerance) as the unfused model."""
    import copy
    from transformers import BertConfig, BertForSequenceClassification
    torch.manual_seed(2)
    cfg = BertConfig(num_labels=2, num_hidden_layers=1)
    base = bf.to_bayesian(BertForSequenceClassification(cfg), delta=0.05, freeze=True, gemm_dtype="bf16")
    fused = bf.accelerate_host_(copy.deepcopy(base), layernorm=False, fuse_gelu=True)
    inter = fused.model.bert.encoder.layer[0].intermediate
    assert inter.dense.activation == "gelu" and isinstance(inter.intermediate_act_fn, torch.nn.Identity)
    S, B, Tn = 2, 8, 128
    ids = torch.randint(0, cfg.vocab_size, (B, Tn), generator=torch.Generator().manual_seed(3)).to(DEV)
    outs = []
    for m in (base, fused):
        m = m.to(DEV).eval()
        bf.cast_frequentist_(m, torch.bfloat16)
        gen = torch.Generator().manual_seed(5)
        for l in m.bayesian_children:
            l.weight.normal = FixedEps([torch.randn(l.weight.mu.shape, generator=gen) for _ in range(S)])
            l.bias.normal = FixedEps([torch.randn(l.bias.mu.shape, generator=gen) for _ in range(S)])
        with bf.mc_samples(S):
            outs.append(m(input_ids=ids.repeat(S, 1)).logits.float())
    assert rel_err(outs[1].detach().cpu().numpy(), outs[0].detach().cpu().numpy()) < BF16_TOL

# ------------------------------------------------------------------ fused clip + AdamW (section 8f row 3)
@pytest.mark.parametrize("max_norm", [None, 0.5])
def test_clip_adamw_matches_torch_cpu(max_norm):
    """bf.optim.ClipAdamW == clip_grad_norm_ + torch.optim.AdamW (the reference loop's optimizer step,
    examples/bert_glue.py:240-241) evaluated by torch on the CPU, over several steps and ragged sizes."""
    gen = torch.Generator().manual_seed(17)
    shapes = [(1,), (7,), (33, 5), (128, 64), (16385,), (3, 16384)]
    ref = [torch.nn.Parameter(torch.randn(sh, generator=gen)) for sh in shapes]
    mine = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ref]
    o_ref = torch.optim.AdamW(ref, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.1)
    o_mine = bf.optim.ClipAdamW(mine, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.1, max_grad_norm=max_norm)
    for it in range(4):
        grads = [torch.randn(sh, generator=gen) * (3.0 if it % 2 else 0.01) for sh in shapes]
        for p, q, g in zip(ref, mine, grads):
            p.grad, q.grad = g.clone(), g.clone().to(DEV)
        if it == 2:  # a tensor without a gradient is skipped by both
            ref[1].grad, mine[1].grad = None, None
        norm_ref = (torch.nn.utils.clip_grad_norm_(ref, max_norm) if max_norm is not None
                    else torch.sqrt(sum((p.grad.double() ** 2).sum() for p in ref if p.grad is not None)))
        o_ref.step()
        norm = o_mine.step()
        assert abs(float(norm) - float(norm_ref)) <= 1e-5 * float(norm_ref)
        for p, q in zip(ref, mine):
            assert rel_err(q.detach().cpu().numpy(), p.detach().numpy()) < 2e-6
    for i, p in enumerate(ref):
        st = o_ref.state[p]
        assert rel_err(o_mine.exp_avg[i].cpu().numpy(), st["exp_avg"].numpy()) < 1e-5
        assert rel_err(o_mine.exp_avg_sq[i].cpu().numpy(), st["exp_avg_sq"].numpy()) < 1e-5


def test_clip_adamw_bf16_params_and_graph_capture():
    """bf16 parameters (fp32 moments) follow the fp32 update to bf16 rounding; step() is capturable."""
    gen = torch.Generator().manual_seed(3)
    w = torch.randn(64, 256, generator=gen)
    ref = torch.nn.Parameter(w.clone())
    q = torch.nn.Parameter(w.clone().to(DEV).bfloat16())
    o_ref = torch.optim.AdamW([ref], lr=1e-2, weight_decay=0.0)
    o = bf.optim.ClipAdamW([q], lr=1e-2, weight_decay=0.0)
    g = torch.randn(64, 256, generator=gen)
    ref.grad = g.bfloat16().float()
    q.grad = g.to(DEV).bfloat16()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        o.step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    ref.data = ref.data.bfloat16().float()
    o_ref.step()
    assert rel_err(q.detach().float().cpu().numpy(), ref.detach().numpy()) < 4e-3
    graph = torch.cuda.CUDAGraph()
    before = q.detach().clone()
    with torch.cuda.graph(graph, stream=side):
        o.step()
    with torch.cuda.stream(side):
        graph.replay()
        graph.replay()
    torch.cuda.synchronize()
    assert float(o.step_count[0]) == 3.0 and not torch.equal(before, q.detach())  # 1 eager + 2 replays

# ------------------------------------------------------------------ CUDA-graph replay == eager (BERT-base layers)
def test_cuda_graph_replay_equals_eager_bert_layers():
    """A captured training step (multi-tensor sampling, folded S-sample fwd, ELBO, bwd) replayed with the
    device-resident step counter at d gives bit-identical logits and rho-gradients to the eager step that
    draws the same Philox stream; a different counter gives a different draw.  BERT-base width, 2 layers."""
    from transformers import BertConfig, BertForSequenceClassification
    torch.manual_seed(0)
    cfg = BertConfig(num_labels=2, num_hidden_layers=2)
    bf.manual_seed(4321)
    bm = bf.to_bayesian(BertForSequenceClassification(cfg), delta=0.05, freeze=True, gemm_dtype="bf16", kl_grad=True)
    bf.accelerate_host_(bm)
    bm = bm.to(DEV).eval()  # eval: dropout off (the layers still sample, quirk Q9)
    bf.enable_presample(bm)
    bf.cast_frequentist_(bm, torch.bfloat16)
    counter = bf.enable_device_step(DEV)
    try:
        S, B, T = 4, 4, 128
        ids = torch.randint(0, cfg.vocab_size, (B, T), generator=torch.Generator().manual_seed(1)).to(DEV)
        rho_params = [l.weight.rho for l in bm.bayesian_children]

        def step():
            for p in bm.parameters():
                p.grad = None
            with bf.mc_samples(S):
                logits = bm(input_ids=ids.repeat(S, 1)).logits.float()
            loss = logits.square().mean() + 1e-6 * (bm.log_variational_posterior() - bm.log_prior()).mean()
            loss.backward()
            return logits

        RUN, D = 7, 3
        # everything runs on one non-default stream: autograd binds each parameter's AccumulateGrad node to the
        # stream of its first use, and the legacy default stream cannot be joined to a capturing stream
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            bm._presampler._runs = RUN - 1
            counter.fill_(D)
            want = step().detach().clone()
            want_g = [p.grad.clone() for p in rho_params]
            for _ in range(2):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in bm.parameters():
            p.grad = None
        bm._presampler._runs = RUN - 1  # the captured launch bakes host step RUN
        g = torch.cuda.CUDAGraph()
        with bf.hf_capture_compat(), torch.cuda.graph(g, stream=side):
            static_logits = step()
        with torch.cuda.stream(side):
            counter.fill_(D)
            g.replay()
        torch.cuda.synchronize()
        assert torch.equal(static_logits, want)
        for p, w in zip(rho_params, want_g):
            # cuDNN's attention backward accumulates dQ with atomics and rounds it to bf16: gradients agree to
            # bf16 noise, not bitwise (the forward above is bit-exact)
            assert rel_err(p.grad.cpu().numpy(), w.cpu().numpy()) < 1e-3
        with torch.cuda.stream(side):
            counter.fill_(D + 1)
            g.replay()
        torch.cuda.synchronize()
        assert not torch.equal(static_logits, want)
    finally:
        bf.disable_device_step()

# ------------------------------------------------------------------ user-loop helpers (section 8f row 4)
def test_sample_bayesian_folded_equals_reference_style_loop():
    """harness.sample_bayesian: one folded forward == the reference's sequential S-loop
    (examples/bert_glue.py:56-73) when both consume the same eps."""
    import copy
    torch.manual_seed(11)
    net = torch.nn.Sequential(torch.nn.Linear(32, 64), torch.nn.ReLU(), torch.nn.Linear(64, 8))
    a = bf.to_bayesian(net, delta=0.05).to(DEV)
    b = copy.deepcopy(a)
    S, B = 5, 6
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(B, 32, generator=gen).to(DEV)
    for m in (a, b):
        g2 = torch.Generator().manual_seed(9)
        for layer in m.bayesian_children:
            layer.weight.normal = FixedEps([torch.randn(layer.weight.mu.shape, generator=g2) for _ in range(S)])
            layer.bias.normal = FixedEps([torch.randn(layer.bias.mu.shape, generator=g2) for _ in range(S)])
    call = lambda out: out  # the model returns the logits tensor itself
    raw_a, mean_a, lp_a, lq_a = bf.sample_bayesian(a, {"input": x}, S, select=call, fold=True)
    raw_b, mean_b, lp_b, lq_b = bf.sample_bayesian(b, {"input": x}, S, select=call, fold=False)
    assert raw_a.shape == (S, B, 8) and mean_a.shape == (B, 8)
    assert rel_err(raw_a.detach().cpu().numpy(), raw_b.detach().cpu().numpy()) < FP32_TOL
    assert abs(float(lp_a) - float(lp_b)) <= FP32_TOL * abs(float(lp_b))
    assert abs(float(lq_a) - float(lq_b)) <= FP32_TOL * abs(float(lq_b))
    stats = bf.predictive_stats(raw_a, torch.zeros(B, dtype=torch.long, device=DEV))
    assert stats["probs"].shape == (B, 8) and float(stats["acc_std"]) >= 0.0

# ------------------------------------------------------------------ multi-tensor sampling (section 8f row 2)
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_presample_equals_per_layer_path(mode):
    """enable_presample: one bf_sample_kl_fwd_multi launch for the whole model gives the same weights,
    log-probs, outputs and gradients as the per-layer kernels fed the same Philox stream."""
    import copy
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(64, 136), torch.nn.Tanh(), torch.nn.Linear(136, 4104, bias=False),
                              torch.nn.Tanh(), torch.nn.Linear(4104, 10))
    bf.manual_seed(77)
    S, B = 3, 8
    x = torch.randn(S * B, 64, device=DEV)
    tol = FP32_TOL if mode == "fp32" else BF16_TOL
    for prior_kw in ({"delta": 0.05}, {}):  # MOPED Gaussian prior / default scale mixture
        a = bf.to_bayesian(copy.deepcopy(net), gemm_dtype=mode, kl_grad=True, **prior_kw).to(DEV)
        b = copy.deepcopy(a)
        bf.enable_presample(a)
        with bf.mc_samples(S):
            ya = a(x)
        la = ya.square().sum() + 1e-3 * (a.log_variational_posterior() - a.log_prior()).sum()
        la.backward()
        # replay the same eps through the per-layer kernels of the twin model
        for la_, lb_ in zip(a.bayesian_children, b.bayesian_children):
            ws, bs = la_._last_streams
            assert ws.step & 0x80000000
            n = lb_.weight.mu.numel()
            lb_.weight.normal = FixedEps([ops.philox_normal(n, ws.seed, ws.step, ws.tensor_id, s, DEV).view_as(lb_.weight.mu)
                                          for s in range(S)])
            if isinstance(lb_.bias, bnn.Gaussian):
                nb = lb_.bias.mu.numel()
                lb_.bias.normal = FixedEps([ops.philox_normal(nb, bs.seed, bs.step, bs.tensor_id, s, DEV) for s in range(S)])
        with bf.mc_samples(S):
            yb = b(x)
        lb = yb.square().sum() + 1e-3 * (b.log_variational_posterior() - b.log_prior()).sum()
        lb.backward()
        assert rel_err(ya.detach().cpu().numpy(), yb.detach().cpu().numpy()) < tol
        assert rel_err(a.log_prior().detach().cpu().numpy(), b.log_prior().detach().cpu().numpy()) < 2e-6
        assert rel_err(a.log_variational_posterior().detach().cpu().numpy(),
                       b.log_variational_posterior().detach().cpu().numpy()) < 2e-6
        for pa, pb in zip(a.parameters(), b.parameters()):
            assert (pa.grad is None) == (pb.grad is None)
            if pa.grad is not None:
                assert rel_err(pa.grad.cpu().numpy(), pb.grad.cpu().numpy()) < tol
        # a second forward draws a different step; a forward without a fresh draw falls back per layer
        with bf.mc_samples(S):
            y2 = a(x)
        assert not torch.equal(y2, ya)

# ------------------------------------------------------------------ native S-sample LayerNorm (row A10)
@pytest.mark.parametrize("S,M,H", [(1, 7, 256), (3, 33, 768), (4, 1000, 768), (2, 129, 1024), (1, 4096, 512)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_native_layernorm_vs_torch_fp64(S, M, H, dtype):
    """bf_layernorm_fwd/bwd through ops.LayerNormFn against F.layer_norm evaluated per
    sample in float64 on the CPU (the torch operator the reference-style layer composes)."""
    gen = torch.Generator().manual_seed(S * 1000 + M + H)
    x = (torch.randn(S * M, H, generator=gen) * 2 + 0.5).to(dtype)
    gamma = 1 + 0.1 * torch.randn(S, H, generator=gen)
    beta = 0.1 * torch.randn(S, H, generator=gen)
    gy = torch.randn(S * M, H, generator=gen).to(dtype)
    xd = x.to(DEV).requires_grad_()
    gd, bd = gamma.to(DEV).requires_grad_(), beta.to(DEV).requires_grad_()
    y = ops.LayerNormFn.apply(xd, gd, bd, S, 1e-12)
    y.backward(gy.to(DEV))
    x64 = x.double().requires_grad_()
    g64, b64 = gamma.double().requires_grad_(), beta.double().requires_grad_()
    y64 = torch.cat([torch.nn.functional.layer_norm(x64[s * M:(s + 1) * M], (H,), g64[s], b64[s], 1e-12) for s in range(S)])
    y64.backward(gy.double())
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    assert y.dtype == dtype
    assert rel_err(y.detach().float().cpu().numpy(), y64.detach().numpy()) < tol
    assert rel_err(xd.grad.float().cpu().numpy(), x64.grad.numpy()) < tol
    # the affine gradients are fp32 sums of fp32 products whatever the activation dtype
    assert rel_err(gd.grad.cpu().numpy(), g64.grad.numpy()) < 2e-5
    assert rel_err(bd.grad.cpu().numpy(), b64.grad.numpy()) < 2e-5
    # deterministic: a second backward gives the same bits
    xd2 = x.to(DEV).requires_grad_()
    gd2, bd2 = gamma.to(DEV).requires_grad_(), beta.to(DEV).requires_grad_()
    ops.LayerNormFn.apply(xd2, gd2, bd2, S, 1e-12).backward(gy.to(DEV))
    assert torch.equal(gd.grad, gd2.grad) and torch.equal(bd.grad, bd2.grad) and torch.equal(xd.grad, xd2.grad)


def test_host_layernorm_shared_affine_matches_torch():
    """accelerate_host_: frequentist nn.LayerNorm through the native kernel == stock module."""
    torch.manual_seed(3)
    ln = torch.nn.LayerNorm(768, eps=1e-12)
    with torch.no_grad():
        ln.weight.add_(0.1 * torch.randn(768)); ln.bias.add_(0.1 * torch.randn(768))
    import copy
    ref = copy.deepcopy(ln).to(DEV)
    fast = bf.accelerate_host_(torch.nn.Sequential(copy.deepcopy(ln))).to(DEV)[0]
    assert type(fast).__name__ == "HostLayerNorm" and list(fast.state_dict()) == list(ref.state_dict())
    x = torch.randn(4, 50, 768, device=DEV)
    xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
    with bf.mc_samples(4):  # folded run: the shared affine must ignore S
        ya = fast(xa)
    yb = ref(xb)
    gy = torch.randn_like(x)
    ya.backward(gy); yb.backward(gy)
    assert rel_err(ya.detach().cpu().numpy(), yb.detach().cpu().numpy()) < FP32_TOL
    assert rel_err(xa.grad.cpu().numpy(), xb.grad.cpu().numpy()) < FP32_TOL
    assert rel_err(fast.weight.grad.cpu().numpy(), ref.weight.grad.cpu().numpy()) < 2e-5
    assert rel_err(fast.bias.grad.cpu().numpy(), ref.bias.grad.cpu().numpy()) < 2e-5

# ------------------------------------------------------------------ Embedding / LayerNorm (rows A9 / A10)
@pytest.mark.parametrize("H", [16, 256])  # 16: torch LayerNorm path, 256: native S-sample LayerNorm kernels
def test_embedding_and_layernorm_against_composed_oracle(H):
    torch.manual_seed(0)
    emb, ln = torch.nn.Embedding(50, H, padding_idx=0), torch.nn.LayerNorm(H)
    with torch.no_grad():
        ln.weight.add_(0.1 * torch.randn(H)); ln.bias.add_(0.1 * torch.randn(H))
    be = bnn.Embedding.from_frequentist(emb, delta=0.05).to(DEV)
    bl = bnn.LayerNorm.from_frequentist(ln, delta=0.05).to(DEV)
    gen = torch.Generator().manual_seed(4)
    S = 3
    e_emb, e_w, e_b = torch.randn(S, 50, H, generator=gen), torch.randn(S, H, generator=gen), torch.randn(S, H, generator=gen)
    be.weight.normal, bl.weight.normal, bl.bias.normal = FixedEps(list(e_emb)), FixedEps(list(e_w)), FixedEps(list(e_b))
    ids = torch.randint(0, 50, (4, 7), generator=gen)
    with bf.mc_samples(S):
        out = bl(be(ids.repeat(S, 1).to(DEV)))
    out.square().sum().backward()
    # oracle: the reference's Gaussian arithmetic composed with F.embedding / F.layer_norm, one sample at a time
    mu_e = emb.weight.detach().clone().requires_grad_(); rho_e = O.moped_rho(emb.weight.detach(), 0.05).requires_grad_()
    mu_w = ln.weight.detach().clone().requires_grad_(); rho_w = O.moped_rho(ln.weight.detach(), 0.05).requires_grad_()
    mu_b = ln.bias.detach().clone().requires_grad_(); rho_b = O.moped_rho(ln.bias.detach(), 0.05).requires_grad_()
    outs, lps, lqs = [], [], []
    for s in range(S):
        We, Ww, Wb = (O.gaussian_sample(m, r, e) for m, r, e in ((mu_e, rho_e, e_emb[s]), (mu_w, rho_w, e_w[s]), (mu_b, rho_b, e_b[s])))
        outs.append(torch.nn.functional.layer_norm(torch.nn.functional.embedding(ids, We, padding_idx=0), (H,), Ww, Wb, ln.eps))
        lqs.append(O.gaussian_log_prob(We.detach(), mu_e.detach(), rho_e.detach()))
        lps.append(O.gaussian_log_prob(We.detach(), emb.weight.detach(), torch.ones(50, H)))
    ref = torch.cat(outs)
    ref.square().sum().backward()
    assert rel_err(out.detach().cpu().numpy(), ref.detach().numpy()) < FP32_TOL
    assert rel_err(be.log_variational_posterior_samples.cpu().numpy(), torch.stack(lqs).numpy()) < FP32_TOL
    assert rel_err(be.log_prior_samples.cpu().numpy(), torch.stack(lps).numpy()) < FP32_TOL
    # the registered scalars stay 0-dim (reference state_dict shapes): mean over the samples
    assert be.log_prior.dim() == 0 and abs(float(be.log_prior) - float(torch.stack(lps).mean())) <= FP32_TOL * abs(float(torch.stack(lps).mean()))
    assert rel_err(be.weight.rho.grad.cpu().numpy(), rho_e.grad.numpy()) < 1e-4
    assert rel_err(bl.weight.rho.grad.cpu().numpy(), rho_w.grad.numpy()) < 1e-4
    assert rel_err(bl.bias.mu.grad.cpu().numpy(), mu_b.grad.numpy()) < 1e-4


@pytest.mark.parametrize("S,M,N,K", [(1, 128, 256, 64), (2, 256, 512, 264), (3, 200, 136, 768), (4, 2560, 768, 768)])
def test_dgrad_accumulate_adds_in_place(S, M, N, K):
    """bf_linear_dgrad_accumulate: dx += gy . w through the TMA reduce-add epilogue (single-CTA and CTA-pair kernels,
    ragged tiles included) against float64."""
    lib = _lib.load()
    gen = torch.Generator().manual_seed(S * 100 + M + N + K)
    gy = torch.randn(S, M, N, generator=gen).bfloat16().to(DEV)
    w = (torch.randn(S, N, K, generator=gen) * 0.05).bfloat16().to(DEV)
    dx0 = torch.randn(S, M, K, generator=gen).bfloat16().to(DEV)
    dx = dx0.clone()
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.bf_linear_dgrad_accumulate(gy.data_ptr(), w.data_ptr(), dx.data_ptr(), S, M, N, K, BF_BF16, BF_BF16, st)
    _lib.check(rc, "bf_linear_dgrad_accumulate")
    want = dx0.double() + torch.einsum("smn,snk->smk", gy.double(), w.double())
    assert rel_err(dx.double().cpu().numpy(), want.cpu().numpy()) < BF16_TOL
    # twice more: still exactly one add per element and launch (no split of the reduction across CTAs)
    for _ in range(2):
        _lib.check(lib.bf_linear_dgrad_accumulate(gy.data_ptr(), w.data_ptr(), dx.data_ptr(), S, M, N, K, BF_BF16, BF_BF16, st), "acc")
    want3 = dx0.double() + 3 * torch.einsum("smn,snk->smk", gy.double(), w.double())
    assert rel_err(dx.double().cpu().numpy(), want3.cpu().numpy()) < 2 * BF16_TOL
    # fp32 mode is refused (the parity path keeps autograd's own accumulation)
    assert lib.bf_linear_dgrad_accumulate(gy.data_ptr(), w.data_ptr(), dx.data_ptr(), S, M, N, K, BF_F32, BF_F32, st) != 0


# ------------------------------------------------------------------ fused dropout + residual + LayerNorm (output blocks)
def test_dropout_mask_matches_oracle_contract():
    for n, p, seed, step, site in [(8 * 1000, 0.1, 1234, 0, 1), (4099, 0.5, 0xDEADBEEFCAFEF00D, 9, 77), (5, 0.25, 3, 2, 1)]:
        got = ops.dropout_mask(n, ops.DropoutSpec(p, seed, site, step), DEV).cpu().numpy()
        assert np.array_equal(got, P.dropout_keep_mask(n, p, seed, step, site))  # integer work: bit-exact
    n = 1 << 22
    m = ops.dropout_mask(n, ops.DropoutSpec(0.1, 5, 3, 1), DEV).float()
    assert abs(float(m.mean()) - 0.9) < 5 * np.sqrt(0.09 / n) + 1e-5
    m2 = ops.dropout_mask(n, ops.DropoutSpec(0.1, 5, 3, 2), DEV).float()  # next step: independent mask
    assert abs(float((m * m2).mean()) - 0.81) < 1e-3
    assert bool(ops.dropout_mask(1000, ops.DropoutSpec(0.0, 5, 3, 1), DEV).all())


@pytest.mark.parametrize("S,M,H,shared", [(1, 7, 256, True), (3, 33, 768, False), (4, 1000, 768, True),
                                          (2, 129, 1024, False), (4, 515, 1024, True), (1, 4096, 512, True)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_resln_vs_torch_fp64(S, M, H, shared, dtype, p):
    """bf_resln_fwd/bwd through ops.ResidualLayerNormFn against float64 torch:
    y = layer_norm(h * keep / (1 - p) + r) per sample, with the keep mask the kernel regenerates."""
    gen = torch.Generator().manual_seed(S * 1000 + M + H)
    h = (torch.randn(S * M, H, generator=gen) * 1.5).to(dtype)
    r = (torch.randn(S * M, H, generator=gen) + 0.3).to(dtype)
    gamma = 1 + 0.1 * torch.randn(H if shared else (S, H), generator=gen)
    beta = 0.1 * torch.randn(H if shared else (S, H), generator=gen)
    gy = torch.randn(S * M, H, generator=gen).to(dtype)
    spec = ops.DropoutSpec(p=p, seed=99, site_id=11, step=4)
    hd, rd = h.to(DEV).requires_grad_(), r.to(DEV).requires_grad_()
    gd, bd = gamma.to(DEV).requires_grad_(), beta.to(DEV).requires_grad_()
    box = []
    y = ops.ResidualLayerNormFn.apply(hd, rd, gd, bd, S, 1e-12, spec, box)
    y.backward(gy.to(DEV))
    keep = ops.dropout_mask(S * M * H, spec, DEV).cpu().view(S * M, H).double()
    h64, r64 = h.double().requires_grad_(), r.double().requires_grad_()
    g64, b64 = gamma.double().requires_grad_(), beta.double().requires_grad_()
    z64 = h64 * keep / (1.0 - np.float32(p).astype(np.float64)) + r64
    y64 = torch.cat([torch.nn.functional.layer_norm(z64[s * M:(s + 1) * M], (H,), g64 if shared else g64[s],
                                                    b64 if shared else b64[s], 1e-12) for s in range(S)])
    y64.backward(gy.double())
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    assert y.dtype == dtype
    assert rel_err(y.detach().float().cpu().numpy(), y64.detach().numpy()) < tol
    assert rel_err(hd.grad.float().cpu().numpy(), h64.grad.numpy()) < tol
    assert rel_err(rd.grad.float().cpu().numpy(), r64.grad.numpy()) < tol
    atol = 2e-5 if dtype == torch.float32 else 1e-2  # bf16: z is rounded before it is normalised
    assert rel_err(gd.grad.cpu().numpy(), g64.grad.numpy()) < atol
    assert rel_err(bd.grad.cpu().numpy(), b64.grad.numpy()) < 2e-5
    # the bias gradient handed to the producing Linear: per-sample column sums of dh
    assert len(box) == 1 and tuple(box[0].shape) == (S, H)
    assert rel_err(box[0].cpu().numpy(), h64.grad.view(S, M, H).sum(1).numpy()) < (2e-5 if dtype == torch.float32 else 2e-3)
    if p > 0:  # dropped positions carry exactly zero gradient
        assert bool((hd.grad.cpu()[keep == 0] == 0).all())
    # deterministic: a second run gives the same bits
    hd2, rd2 = h.to(DEV).requires_grad_(), r.to(DEV).requires_grad_()
    gd2, bd2 = gamma.to(DEV).requires_grad_(), beta.to(DEV).requires_grad_()
    box2 = []
    y2 = ops.ResidualLayerNormFn.apply(hd2, rd2, gd2, bd2, S, 1e-12, spec, box2)
    y2.backward(gy.to(DEV))
    assert torch.equal(y, y2) and torch.equal(hd.grad, hd2.grad) and torch.equal(rd.grad, rd2.grad)
    assert torch.equal(gd.grad, gd2.grad) and torch.equal(bd.grad, bd2.grad) and torch.equal(box[0], box2[0])
    if p > 0:
        # the backward normally LOADS the keep bits forward stored (128 bytes per row); regenerating them from the Philox
        # counter instead (ops.resln_keep_bits off) gives the same bits everywhere
        ops.resln_keep_bits["on"] = False
        try:
            hd3, rd3 = h.to(DEV).requires_grad_(), r.to(DEV).requires_grad_()
            gd3, bd3 = gamma.to(DEV).requires_grad_(), beta.to(DEV).requires_grad_()
            box3 = []
            y3 = ops.ResidualLayerNormFn.apply(hd3, rd3, gd3, bd3, S, 1e-12, spec, box3)
            y3.backward(gy.to(DEV))
        finally:
            ops.resln_keep_bits["on"] = True
        assert torch.equal(y, y3) and torch.equal(hd.grad, hd3.grad) and torch.equal(rd.grad, rd3.grad)
        assert torch.equal(gd.grad, gd3.grad) and torch.equal(bd.grad, bd3.grad) and torch.equal(box[0], box3[0])


@pytest.mark.parametrize("mode,bayes_ln,sinks", [("fp32", False, False), ("bf16", False, False), ("fp32", True, False),
                                                 ("bf16", False, True), ("bf16", True, True)])
def test_accelerate_host_fuses_hf_output_blocks(mode, bayes_ln, sinks):
    """accelerate_host_(fuse_residual=True) on HF BERT layers: with dropout off the fused model reproduces the
    unfused one (logits, log-probs and every gradient incl. the bias gradients handed over by the fused
    backward); with dropout on, the block equals its manual composition under the regenerated mask."""
    import copy
    from transformers import BertConfig, BertForSequenceClassification
    torch.manual_seed(2)
    cfg = BertConfig(num_labels=2, num_hidden_layers=2, hidden_size=256, intermediate_size=512, num_attention_heads=4)
    layers = bnn.TORCH2BAYE_ALL if bayes_ln else None
    base = bf.to_bayesian(BertForSequenceClassification(cfg), delta=0.05, freeze=False, gemm_dtype=mode, kl_grad=True,
                          layers=layers)
    fused = bf.accelerate_host_(copy.deepcopy(base), layernorm=False, fuse_gelu=False, fuse_residual=True,
                                grad_sinks=sinks)
    blocks = [m for m in fused.modules() if type(m).__name__.startswith("Fused")]
    assert len(blocks) == 4 and list(fused.state_dict()) == list(base.state_dict())
    S, B, Tn = 2, 4, 16
    ids = torch.randint(0, cfg.vocab_size, (B, Tn), generator=torch.Generator().manual_seed(3)).to(DEV)
    outs = []
    acc0 = ops.stats["dgrad_accumulated"]
    for m in (base, fused):
        bf.runtime.enable_grad_sinks(sinks and m is fused)
        m = m.to(DEV).eval()  # eval(): dropout off, so both models compute the same function
        if mode == "bf16":
            bf.cast_frequentist_(m, torch.bfloat16)
        gen = torch.Generator().manual_seed(5)
        for l in m.bayesian_children:
            for g in (l.weight, getattr(l, "bias", None)):
                if isinstance(g, bnn.Gaussian):
                    g.normal = FixedEps([torch.randn(g.mu.shape, generator=gen) for _ in range(S)])
        with bf.mc_samples(S):
            logits = m(input_ids=ids.repeat(S, 1)).logits.float()
        lp, lq = m.log_prior(), m.log_variational_posterior()
        (logits.square().sum() + 1e-3 * (lq - lp).sum()).backward()
        outs.append((logits.detach(), lp.detach(), lq.detach(), {n: p.grad for n, p in m.named_parameters() if p.grad is not None}))
    bf.runtime.enable_grad_sinks(False)
    # with sinks, every layer input that feeds Linear layers and a fused residual took the in-place path:
    # 2 layers x (q, k, v, intermediate.dense) accumulating dgrads
    assert ops.stats["dgrad_accumulated"] - acc0 == (8 if sinks else 0)
    tol = FP32_TOL * 10 if mode == "fp32" else BF16_TOL  # fp32: different (still fp32) summation orders through 2 layers
    assert rel_err(outs[1][0].cpu().numpy(), outs[0][0].cpu().numpy()) < tol
    assert rel_err(outs[1][1].cpu().numpy(), outs[0][1].cpu().numpy()) < FP32_TOL
    assert rel_err(outs[1][2].cpu().numpy(), outs[0][2].cpu().numpy()) < FP32_TOL
    assert outs[0][3].keys() == outs[1][3].keys()
    gtol = 1e-3 if mode == "fp32" else 5e-2
    for n in outs[0][3]:
        assert rel_err(outs[1][3][n].float().cpu().numpy(), outs[0][3][n].float().cpu().numpy()) < gtol, n
    # dropout on: one block against its manual composition with the mask the kernels regenerate
    blk = blocks[1].train()
    H = cfg.hidden_size
    dt = torch.bfloat16 if mode == "bf16" else torch.float32
    x = torch.randn(S * B * Tn, cfg.intermediate_size, device=DEV, dtype=dt)
    res = torch.randn(S * B * Tn, H, device=DEV, dtype=dt)
    gen = torch.Generator().manual_seed(6)
    eps_w = [torch.randn(blk.dense.weight.mu.shape, generator=gen) for _ in range(S)]
    eps_b = [torch.randn(blk.dense.bias.mu.shape, generator=gen) for _ in range(S)]
    ln = blk.LayerNorm
    eps_ln = [[torch.randn(H, generator=gen) for _ in range(S)] for _ in range(2)] if bayes_ln else None

    def arm():
        blk.dense.weight.normal, blk.dense.bias.normal = FixedEps(list(eps_w)), FixedEps(list(eps_b))
        if bayes_ln:
            ln.weight.normal, ln.bias.normal = FixedEps(list(eps_ln[0])), FixedEps(list(eps_ln[1]))

    arm()
    with bf.mc_samples(S):
        y = blk(x, res)
    spec = blk._last_dropout
    assert spec.p == pytest.approx(cfg.hidden_dropout_prob)
    keep = ops.dropout_mask(y.numel(), spec, DEV).view_as(y)
    arm()
    with bf.mc_samples(S):
        h = blk.dense(x)
        want = ln(h * keep.to(h.dtype) / (1 - spec.p) + res)
    assert rel_err(y.detach().float().cpu().numpy(), want.detach().float().cpu().numpy()) < (FP32_TOL if mode == "fp32" else BF16_TOL)


# ------------------------------------------------------------------ BASELINE-size properties (config 2: 4096x4096)
def test_full_size_properties_config2():
    n, S = 4096 * 4096, 4
    gen = torch.Generator(device=DEV).manual_seed(0)
    mu = torch.empty(n, device=DEV).uniform_(-0.2, 0.2, generator=gen)
    rho = torch.empty(n, device=DEV).uniform_(-5, -4, generator=gen)
    pr = ops.PriorSpec(BF_PRIOR_MIXTURE, 0.5, 1.0, float(np.float32(np.exp(-6))))
    w, lq, lp = run_sample_kl(mu, rho, pr, S, seed=7, tid=3)
    # additivity over a split of the tensor is NOT expected bitwise, but to fp32 sum accuracy
    h = n // 2
    # the second half of the tensor uses quads offset by h/4: reproduce through injected eps from the stream
    eps = torch.stack([ops.philox_normal(n, 7, 0, 3, s, DEV) for s in range(S)])
    _, lq_a, lp_a = run_sample_kl(mu[:h], rho[:h], pr, S, eps[:, :h].contiguous())
    _, lq_b, lp_b = run_sample_kl(mu[h:], rho[h:], pr, S, eps[:, h:].contiguous())
    assert torch.allclose(lq, lq_a + lq_b, rtol=2e-6) and torch.allclose(lp, lp_a + lp_b, rtol=2e-6)
    # float64 evaluation of the same contract on the device
    sig = torch.nn.functional.softplus(rho.double())
    for s in range(S):
        lq64 = (-0.5 * np.log(2 * np.pi) - sig.log() - 0.5 * ((w[s].double() - mu.double()) / sig) ** 2).sum()
        assert abs(float(lq[s]) - float(lq64)) <= FP32_TOL * abs(float(lq64))
    # sample statistics of the standardised draw
    z = ((w[0].double() - mu.double()) / sig)
    assert abs(float(z.mean())) < 5 / np.sqrt(n) and abs(float(z.var()) - 1) < 5 * np.sqrt(2 / n)


def test_row_kernels_take_misaligned_views():
    """A dense view whose storage offset breaks 16-byte alignment (slice of a flat buffer) is copied by the binding
    instead of reaching the vector / bulk-copy kernels; the C ABI itself refuses such pointers."""
    H, rows = 256, 64
    flat = torch.randn(rows * H + 8, device=DEV).bfloat16()
    h = flat[1:1 + rows * H].view(rows, H)           # 2-byte offset
    assert h.data_ptr() % 16 != 0 and h.is_contiguous()
    r = torch.randn(rows, H, device=DEV).bfloat16().requires_grad_()
    g, b = torch.ones(H, device=DEV, requires_grad=True), torch.zeros(H, device=DEV, requires_grad=True)
    y = ops.ResidualLayerNormFn.apply(h, r, g, b, 1, 1e-5, ops.DropoutSpec(), None)
    want = torch.nn.functional.layer_norm((h.float() + r.float()), (H,), g, b, 1e-5)
    assert rel_err(y.float().detach().cpu().numpy(), want.detach().cpu().numpy()) < BF16_TOL
    gy = torch.randn(rows * H + 8, device=DEV).bfloat16()[1:1 + rows * H].view(rows, H)
    y.backward(gy)
    assert r.grad is not None and torch.isfinite(r.grad.float()).all()
    y2 = ops.LayerNormFn.apply(h, g, b, 1, 1e-5)
    assert rel_err(y2.float().detach().cpu().numpy(),
                   torch.nn.functional.layer_norm(h.float(), (H,), g, b, 1e-5).detach().cpu().numpy()) < BF16_TOL
    lib = _lib.load()
    z = torch.empty(rows, H, device=DEV, dtype=torch.bfloat16)
    mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    rc = lib.bf_resln_fwd(h.data_ptr(), r.data_ptr(), BF_BF16, g.data_ptr(), b.data_ptr(), 0, 1, rows, H, 1e-5, 0.0, 1, 0, 1,
                          z.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc != 0 and b"16-byte aligned" in lib.bf_last_error()


# ------------------------------------------------------------------ reference precision on the tensor cores (fp32x3)
@pytest.mark.parametrize("S,M,N,K", [(1, 128, 256, 64), (2, 256, 512, 256), (3, 200, 136, 264), (2, 1000, 768, 3072),
                                     (4, 4096, 768, 768)])
def test_x3_contractions_vs_float64(S, M, N, K):
    """bf_linear_{fwd,dgrad,wgrad}_x3: fp32 operands split into bf16 (hi, lo) pairs, three tcgen05 passes per tile;
    results must sit within 1e-5 (norm-wise) of the float64 contraction of the fp32 operands -- the tolerance the north
    star gives the fp32 mode -- where plain bf16 operands would be ~3e-3 off."""
    lib = _lib.load()
    gen = torch.Generator().manual_seed(S * 7 + M + K)
    x = torch.randn(S, M, K, generator=gen).to(DEV)
    w = (torch.randn(S, N, K, generator=gen) * 0.05).to(DEV)
    b = torch.randn(S, N, generator=gen).to(DEV)
    gy = torch.randn(S, M, N, generator=gen).to(DEV)
    xs, ws, gs = ops.split_bf16x2(x), ops.split_bf16x2(w), ops.split_bf16x2(gy)
    assert rel_err((xs[0].double() + xs[1].double()).cpu().numpy(), x.double().cpu().numpy()) < 2 ** -16
    st = torch.cuda.current_stream().cuda_stream
    y = torch.empty(S, M, N, device=DEV)
    _lib.check(lib.bf_linear_fwd_x3(xs[0].data_ptr(), xs[1].data_ptr(), ws[0].data_ptr(), ws[1].data_ptr(), b.data_ptr(),
                                    y.data_ptr(), S, M, N, K, st), "fwd_x3")
    want = torch.einsum("smk,snk->smn", x.double(), w.double()) + b.double()[:, None, :]
    e_fwd = rel_err(y.cpu().numpy(), want.cpu().numpy())
    dx = torch.empty(S, M, K, device=DEV)
    _lib.check(lib.bf_linear_dgrad_x3(gs[0].data_ptr(), gs[1].data_ptr(), ws[0].data_ptr(), ws[1].data_ptr(), dx.data_ptr(),
                                      S, M, N, K, st), "dgrad_x3")
    e_dg = rel_err(dx.cpu().numpy(), torch.einsum("smn,snk->smk", gy.double(), w.double()).cpu().numpy())
    dw = torch.empty(S, N, K, device=DEV)
    _lib.check(lib.bf_linear_wgrad_x3(gs[0].data_ptr(), gs[1].data_ptr(), xs[0].data_ptr(), xs[1].data_ptr(), dw.data_ptr(),
                                      S, M, N, K, st), "wgrad_x3")
    e_wg = rel_err(dw.cpu().numpy(), torch.einsum("smn,smk->snk", gy.double(), x.double()).cpu().numpy())
    print(f"[x3 {S}x{M}x{N}x{K}] fwd {e_fwd:.2e} dgrad {e_dg:.2e} wgrad {e_wg:.2e}")
    assert max(e_fwd, e_dg, e_wg) < FP32_TOL


@pytest.mark.parametrize("tag", ["default", "nobias", "moped_tc", "moped_frozen_tc", "moped_frozen_tiles"])
def test_linear_fp32x3_matches_reference(tag):
    """gemm_dtype="fp32x3" against outputs of the unmodified reference at the fp32 tolerance (1e-5), with every
    contraction on the tensor cores."""
    g = load_golden("linear.npz")
    layer, bias, freeze = _layer_from_golden(g, tag, "fp32x3")
    x = T(g[f"{tag}_x"]).requires_grad_()
    ops.enable_kernel_timing(True)
    try:
        y = layer(x)
        y.backward(T(g[f"{tag}_gy"]))
        torch.cuda.synchronize()
        ran = set(ops.kernel_timing_summary())
    finally:
        ops.enable_kernel_timing(False)
    assert {"gemm_fwd_x3", "gemm_dgrad_x3", "gemm_wgrad_x3"} <= ran, ran
    assert rel_err(y.detach().cpu().numpy(), g[f"{tag}_y"]) < FP32_TOL
    assert abs(float(layer.log_prior) - float(g[f"{tag}_log_prior"])) <= FP32_TOL * abs(float(g[f"{tag}_log_prior"]))
    assert rel_err(x.grad.cpu().numpy(), g[f"{tag}_g_x"]) < FP32_TOL
    assert rel_err(layer.weight.rho.grad.cpu().numpy(), g[f"{tag}_g_w_rho"]) < 2 * FP32_TOL
    if not freeze:
        assert rel_err(layer.weight.mu.grad.cpu().numpy(), g[f"{tag}_g_w_mu"]) < FP32_TOL
    if bias:
        assert rel_err(layer.bias.rho.grad.cpu().numpy(), g[f"{tag}_g_b_rho"]) < FP32_TOL


def test_tiny_bert_fp32x3_mode_within_fp32_tolerance():
    g = load_golden("tiny_bert.npz")
    S = int(g["S"])
    bm, layers = _tiny_bert(g, "fp32x3")
    for i, l in enumerate(layers):
        l.weight.normal = FixedEps(list(g[f"eps_w{i}"]))
        l.bias.normal = FixedEps(list(g[f"eps_b{i}"]))
    with bf.mc_samples(S):
        out = bm(input_ids=T(g["ids"]).repeat(S, 1)).logits
    raw = out.view(S, -1, out.shape[-1])
    assert rel_err(raw.detach().float().cpu().numpy(), g["logits"]) < 5 * FP32_TOL
    assert rel_err(bm.log_prior().cpu().numpy(), g["log_prior"]) < FP32_TOL


# ------------------------------------------------------------------ GELU' folded into the consumer's dgrad
@pytest.mark.parametrize("S,M,N,K", [(2, 2560, 768, 3072), (4, 2600, 264, 520), (1, 20000, 64, 256)])
def test_dgrad_gelu_kernel_vs_float64(S, M, N, K):
    """bf_linear_dgrad_gelu: gz = (gy . w) o gelu'(z) from the dgrad epilogue (z tile TMA-loaded next to the operands)
    against float64; ragged tiles and a tile count that is not a multiple of the SM-pair count included."""
    lib = _lib.load()
    if not lib.bf_linear_dgrad_gelu_supported(S, M, N, K):
        pytest.skip("shape below the CTA-pair kernel's occupancy threshold")
    gen = torch.Generator().manual_seed(S + M + N + K)
    gy = torch.randn(S, M, N, generator=gen).bfloat16().to(DEV)
    w = (torch.randn(S, N, K, generator=gen) * 0.05).bfloat16().to(DEV)
    z = (torch.randn(S, M, K, generator=gen) * 1.5).bfloat16().to(DEV)
    gz = torch.full((S, M, K), float("nan"), dtype=torch.bfloat16, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.bf_linear_dgrad_gelu(gy.data_ptr(), w.data_ptr(), z.data_ptr(), gz.data_ptr(), S, M, N, K, st), "dgrad_gelu")
    zd = z.double()
    dgelu = 0.5 * (1 + torch.erf(zd / np.sqrt(2))) + zd * torch.exp(-0.5 * zd * zd) / np.sqrt(2 * np.pi)
    want = torch.einsum("smn,snk->smk", gy.double(), w.double()) * dgelu
    assert torch.isfinite(gz.float()).all()
    assert rel_err(gz.float().cpu().numpy(), want.cpu().numpy()) < 4e-3  # bf16 rounding of the result
    # bit-identical to the unfused composition's arithmetic up to the intermediate bf16 rounding of dx
    dx = torch.empty(S, M, K, dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.bf_linear_dgrad(gy.data_ptr(), w.data_ptr(), dx.data_ptr(), S, M, N, K, BF_BF16, BF_BF16, st), "dgrad")
    assert rel_err(gz.float().cpu().numpy(), (dx.double() * dgelu).cpu().numpy()) < 6e-3
    # the same launch with the column sums of gz from its epilogue (per-block partial rows + fixed-order reduction):
    # identical gz, sums of the bf16-ROUNDED values it stored, bit-identical from run to run
    gz2, db = torch.empty_like(gz), torch.full((S, K), float("nan"), device=DEV)
    ws = torch.zeros(int(lib.bf_linear_dgrad_gelu_bias_workspace_bytes(S, M, K)), dtype=torch.uint8, device=DEV)
    for rep in range(2):
        _lib.check(lib.bf_linear_dgrad_gelu_bias(gy.data_ptr(), w.data_ptr(), z.data_ptr(), gz2.data_ptr(), db.data_ptr(),
                                                 ws.data_ptr(), S, M, N, K, st), "dgrad_gelu_bias")
        torch.cuda.synchronize()
        if rep == 0:
            db0 = db.clone()
    assert torch.equal(gz2, gz) and torch.equal(db, db0)
    want_db = gz.double().sum(1)
    assert rel_err(db.cpu().numpy(), want_db.cpu().numpy()) < 1e-5


def test_ffn_block_with_and_without_gelu_link():
    """up(gelu) -> down: with the GeluLink the down layer's dgrad applies gelu'(z) and the up layer skips its GELU' pass;
    outputs are identical and every gradient agrees with the unlinked path to bf16 rounding.  A second consumer of the
    activation must be detected."""
    import copy
    torch.manual_seed(3)
    H, F_, S, B = 768, 3072, 2, 1280
    up, down = torch.nn.Linear(H, F_), torch.nn.Linear(F_, H)
    bf.manual_seed(5)
    a_up = bnn.Linear.from_frequentist(up, delta=0.05, freeze=True).to(DEV)
    a_down = bnn.Linear.from_frequentist(down, delta=0.05, freeze=True).to(DEV)
    for l in (a_up, a_down):
        l.gemm_dtype = torch.bfloat16
    a_up.activation = "gelu"
    x = torch.randn(S * B, H, device=DEV).bfloat16()
    gy = torch.randn(S * B, H, device=DEV).bfloat16()
    results = {}
    for linked in (True, False):
        bf.runtime.enable_gelu_links(linked)
        try:
            for l in (a_up, a_down):
                l.weight.step = l.bias.step = 0
                l.zero_grad(set_to_none=True)
            xa = x.clone().requires_grad_()
            ops.enable_kernel_timing(True)
            with bf.mc_samples(S):
                a = a_up(xa)
                assert (getattr(a, "_bf_gelu_link", None) is not None) == linked
                y = a_down(a)
            y.backward(gy)
            torch.cuda.synchronize()
            ran = ops.kernel_timing_summary()
            ops.enable_kernel_timing(False)
            assert ("gelu_bwd_bias_grad" in ran) == (not linked)
            # linked: the up layer's bias gradient came out of the down layer's dgrad epilogue -- the only separate
            # bias-gradient pass left is the down layer's own
            assert ran["bias_grad"]["calls"] == 1
            results[linked] = (y.detach().clone(), xa.grad.clone(), a_up.weight.rho.grad.clone(), a_up.bias.rho.grad.clone(),
                               a_down.weight.rho.grad.clone())
        finally:
            ops.enable_kernel_timing(False)
            bf.runtime.enable_gelu_links(True)
    assert torch.equal(results[True][0], results[False][0])
    for got, want in zip(results[True][1:], results[False][1:]):
        assert rel_err(got.float().cpu().numpy(), want.float().cpu().numpy()) < 6e-3
    # a second consumer of the activation: the hook must refuse
    for l in (a_up, a_down):
        l.zero_grad(set_to_none=True)
    xa = x.clone().requires_grad_()
    with bf.mc_samples(S):
        a = a_up(xa)
        y = a_down(a)
    with pytest.raises(RuntimeError, match="fused GELU"):
        (y.float().sum() + a.float().sum()).backward()


# ------------------------------------------------------------------ native short-sequence attention (host model, opt-in)
@pytest.mark.parametrize("B,H,Tn", [(3, 12, 128), (2, 4, 64), (5, 2, 16), (2, 3, 112), (150, 2, 128)])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_attention_kernels_vs_float64(B, H, Tn, p):
    """bf_attention_fwd / _bwd (whole sequence per block, mma.sync, Philox keep mask regenerated in backward) against a
    float64 composition softmax(q k^T * scale) -> mask / keep-probability -> . v and its autograd.  q, k, v are the
    transposed views of [B, T, H*64] buffers HuggingFace hands its attention function."""
    Dh = 64
    gen = torch.Generator().manual_seed(B * 1000 + H * 10 + Tn)
    bufs = [(torch.randn(B, Tn, H * Dh, generator=gen) * 0.8).bfloat16().to(DEV).requires_grad_() for _ in range(3)]
    q, k, v = (t.view(B, Tn, H, Dh).transpose(1, 2) for t in bufs)
    drop = ops.DropoutSpec(p=p, seed=0xABCDEF0123, site_id=7, step=3)
    scale = Dh ** -0.5
    out = ops.AttentionFn.apply(q, k, v, scale, drop)
    assert out.shape == (B, Tn, H, Dh) and out.dtype == torch.bfloat16
    gout = torch.randn(out.shape, generator=gen).bfloat16().to(DEV)
    out.backward(gout)
    keep = ops.attention_dropout_mask(B, H, Tn, drop, DEV).double()
    if p == 0.0:
        assert bool((keep == 1).all())
        keep_prob = 1.0
    else:
        keep_prob = 1.0 - round(p * 65536) / 65536
        assert abs(float(keep.mean()) - keep_prob) < 5.0 * np.sqrt(p * (1 - p) / keep.numel()) + 1e-4
    qd, kd, vd = (t.detach().double().view(B, Tn, H, Dh).transpose(1, 2).requires_grad_() for t in bufs)
    P = torch.softmax(qd @ kd.transpose(-1, -2) * scale, dim=-1)
    ref = ((P * keep / keep_prob) @ vd).transpose(1, 2)  # [B, T, H, D]
    ref.backward(gout.double())
    assert rel_err(out.detach().float().cpu().numpy(), ref.detach().cpu().numpy()) < 6e-3
    for got, want, name in zip(bufs, (qd, kd, vd), "qkv"):
        g = got.grad.view(B, Tn, H, Dh).transpose(1, 2)
        e = rel_err(g.float().cpu().numpy(), want.grad.cpu().numpy())
        assert e < 1.5e-2, (name, e)
    # deterministic: a second run gives the same bits
    for t in bufs:
        t.grad = None
    out2 = ops.AttentionFn.apply(q, k, v, scale, drop)
    out2.backward(gout)
    assert torch.equal(out, out2)
    g1 = bufs[0].grad.clone()
    for t in bufs:
        t.grad = None
    ops.AttentionFn.apply(q, k, v, scale, drop).backward(gout)
    assert torch.equal(g1, bufs[0].grad)


def test_attention_tcgen05_and_mma_sync_kernels_share_one_mask_and_agree():
    """T = 128 has two kernel families (tcgen05: bf_attention_tc.cu, default; mma.sync: bf_attention.cu,
    BF_OPT_ATTN_TC = 0).  The keep bits the tcgen05 forward stores for its backward are BIT-IDENTICAL to the mask
    bf_attention_dropout_mask defines (the one the mma.sync kernels regenerate), both families read q / k / v in place
    from a fused [B, T, 3 * H * 64] projection buffer (token stride 3 * H * 64), and their outputs and gradients agree to
    bf16 rounding.  150 x 2 pairs > 148 blocks: the persistent loops wrap."""
    lib = _lib.load()
    B, H, Tn, Dh = 150, 2, 128, 64
    gen = torch.Generator().manual_seed(77)
    fused = (torch.randn(B, Tn, 3 * H * Dh, generator=gen) * 0.7).bfloat16().to(DEV).requires_grad_()
    q, k, v = (fused[..., i * H * Dh:(i + 1) * H * Dh].view(B, Tn, H, Dh).transpose(1, 2) for i in range(3))
    gout = torch.randn(B, Tn, H, Dh, generator=gen).bfloat16().to(DEV)
    drop = ops.DropoutSpec(p=0.25, seed=0x1234567, site_id=9, step=11)
    res = {}
    try:
        for tc in (1, 0):
            lib.bf_set_option(_lib.BF_OPT_ATTN_TC, tc)
            fused.grad = None
            ops.enable_kernel_timing(True)
            out = ops.AttentionFn.apply(q, k, v, Dh ** -0.5, drop)
            saved = out.grad_fn.saved_tensors
            out.backward(gout)
            torch.cuda.synchronize()
            ops.enable_kernel_timing(False)
            res[tc] = (out.detach().float(), fused.grad.detach().float().clone(), saved[-1])
    finally:
        ops.enable_kernel_timing(False)
        lib.bf_set_option(_lib.BF_OPT_ATTN_TC, 1)
    # the stored keep words against the byte mask of the contract
    keep_words = res[1][2]
    assert keep_words is not None and keep_words.shape == (B, H, Tn, 4) and keep_words.dtype == torch.int32
    bits = ((keep_words.view(B, H, Tn, 4, 1) >> torch.arange(32, device=DEV, dtype=torch.int32)) & 1).reshape(B, H, Tn, 128)
    mask = ops.attention_dropout_mask(B, H, Tn, drop, DEV)
    assert torch.equal(bits.to(torch.uint8), mask)
    assert abs(float(mask.float().mean()) - 0.75) < 2e-3
    # the two families against each other
    assert rel_err(res[1][0].cpu().numpy(), res[0][0].cpu().numpy()) < 6e-3
    assert rel_err(res[1][1].cpu().numpy(), res[0][1].cpu().numpy()) < 1.5e-2
    # the tcgen05 backward can also emit the q / k / v projections' bias gradients: column sums of dq, dk, dv per folded
    # sample (S = 3 samples x 50 sequences), equal to the sums of the bf16 gradients it wrote, bit-identical run to run
    S = 3
    got = []
    for rep in range(2):
        boxes = [[], [], []]
        fused.grad = None
        ops.AttentionFn.apply(q, k, v, Dh ** -0.5, drop, boxes, S).backward(gout)
        torch.cuda.synchronize()
        assert all(len(b) == 1 and b[0].shape == (S, H * Dh) and b[0].dtype == torch.float32 for b in boxes)
        got.append(torch.stack([b[0] for b in boxes]))
        want = fused.grad.view(S, B // S * Tn, 3, H * Dh).double().sum(1).transpose(0, 1)  # [3, S, H * Dh]
        assert rel_err(got[-1].cpu().numpy(), want.cpu().numpy()) < 1e-5
    assert torch.equal(got[0], got[1])


@pytest.mark.parametrize("B,H,Tn,p,seed,step,site", [(2, 3, 128, 0.1, 0x1234567, 5, 7), (1, 2, 64, 0.5, 0xDEADBEEFCAFEF00D, 9, 77),
                                                      (3, 1, 16, 0.25, 3, 0, 1)])
def test_attention_dropout_mask_matches_oracle_contract(B, H, Tn, p, seed, step, site):
    """The keep mask of the attention kernels against its CPU restatement (oracle/philox_oracle.attention_keep_mask):
    integer work, bit-exact.  (The tcgen05 forward's stored keep bits are tied to this mask kernel bit for bit in
    test_attention_tcgen05_and_mma_sync_kernels_share_one_mask_and_agree.)"""
    got = ops.attention_dropout_mask(B, H, Tn, ops.DropoutSpec(p, seed, site, step), DEV).cpu().numpy()
    assert np.array_equal(got, P.attention_keep_mask(B, H, Tn, p, seed, step, site))


def test_attention_dropout_masks_are_independent_across_sites_and_steps():
    B, H, Tn = 4, 12, 128
    base = ops.DropoutSpec(p=0.1, seed=11, site_id=3, step=5)
    m0 = ops.attention_dropout_mask(B, H, Tn, base, DEV).float().flatten()
    n = m0.numel()
    for other in (ops.DropoutSpec(0.1, 11, 4, 5), ops.DropoutSpec(0.1, 11, 3, 6), ops.DropoutSpec(0.1, 12, 3, 5)):
        m1 = ops.attention_dropout_mask(B, H, Tn, other, DEV).float().flatten()
        corr = float(torch.corrcoef(torch.stack([m0, m1]))[0, 1])
        assert abs(corr) < 5.0 / np.sqrt(n)
    m = m0.view(B * H * Tn, Tn)
    assert abs(float(torch.corrcoef(torch.stack([m[:, :-1].flatten(), m[:, 1:].flatten()]))[0, 1])) < 5.0 / np.sqrt(n)
    assert torch.equal(m0, ops.attention_dropout_mask(B, H, Tn, base, DEV).float().flatten())


def test_accelerate_host_native_attention_matches_sdpa():
    """accelerate_host_(attention=True) on a HuggingFace BERT: eval-mode logits agree with the torch SDPA path to bf16
    rounding; the rho gradients of every layer are no further from an all-fp32 run of the same model than the bf16 SDPA
    path's are (the key projections' gradients nearly cancel, so two bf16 paths differ by more from each other than
    either does from fp32); the native kernels really ran, and a masked batch falls back to SDPA."""
    import copy
    from transformers import BertConfig, BertForSequenceClassification
    torch.manual_seed(2)
    cfg = BertConfig(vocab_size=100, hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                     intermediate_size=512, max_position_embeddings=128, num_labels=2)
    base = bf.to_bayesian(BertForSequenceClassification(cfg), delta=0.05, freeze=True, gemm_dtype="bf16")
    fast = bf.accelerate_host_(copy.deepcopy(base), layernorm=False, fuse_gelu=False, attention=True, attention_bias_grads=True)
    truth = copy.deepcopy(base)  # fp32 everywhere: FFMA contractions, fp32 host model
    for l in truth.bayesian_children:
        l.gemm_dtype = torch.float32
    base, fast, truth = base.eval().to(DEV), fast.eval().to(DEV), truth.eval().to(DEV)
    for m in (base, fast):
        bf.cast_frequentist_(m, torch.bfloat16)
    S, B, Tn = 2, 3, 128
    ids = torch.randint(0, 100, (B, Tn), generator=torch.Generator().manual_seed(1)).to(DEV)
    gen = torch.Generator().manual_seed(5)
    for la, lb, lc in zip(base.bayesian_children, fast.bayesian_children, truth.bayesian_children):
        ew = [torch.randn(la.weight.mu.shape, generator=gen) for _ in range(S)]
        eb = [torch.randn(la.bias.mu.shape, generator=gen) for _ in range(S)]
        for l in (la, lb, lc):
            l.weight.normal, l.bias.normal = FixedEps([e.clone() for e in ew]), FixedEps([e.clone() for e in eb])
    outs = []
    for m in (base, fast, truth):
        ops.enable_kernel_timing(True)
        try:
            with bf.mc_samples(S):
                logits = m(input_ids=ids.repeat(S, 1)).logits
            logits.float().square().sum().backward()
            torch.cuda.synchronize()
            summary = ops.kernel_timing_summary()
            ran = set(summary)
        finally:
            ops.enable_kernel_timing(False)
        assert ({"attention_fwd", "attention_bwd"} <= ran) == (m is fast)
        if m is not truth:
            # 2 layers x (q, k, v, attention-output, FFN-up, FFN-down) + pooler + classifier = 14 Bayesian Linears; with
            # the native attention the q / k / v bias gradients come out of its backward kernel
            n_bias = summary["bias_grad"]["calls"]
            assert n_bias == (14 - 6 if m is fast else 14), n_bias
        outs.append((logits.detach().float(), [l.weight.rho.grad.clone() for l in m.bayesian_children]))
    assert rel_err(outs[1][0].cpu().numpy(), outs[0][0].cpu().numpy()) < 2e-2
    e_sdpa = [rel_err(g.cpu().numpy(), t.cpu().numpy()) for g, t in zip(outs[0][1], outs[2][1])]
    e_native = [rel_err(g.cpu().numpy(), t.cpu().numpy()) for g, t in zip(outs[1][1], outs[2][1])]
    print(f"[native attention vs SDPA] logits {rel_err(outs[1][0].cpu().numpy(), outs[0][0].cpu().numpy()):.2e}; rho "
          f"grads against fp32: bf16 SDPA max {max(e_sdpa):.2e}, native max {max(e_native):.2e}")
    for a, b in zip(e_native, e_sdpa):
        assert a < max(2.0 * b, 2e-2), (e_native, e_sdpa)
    # an attention mask is outside the kernels' case: transformers' own SDPA function takes it
    mask = torch.ones(B, Tn, dtype=torch.long, device=DEV)
    mask[:, 100:] = 0
    for l in fast.bayesian_children:  # back to the Philox stream
        l.weight.normal = torch.distributions.Normal(l.weight.zero, l.weight.one)
        l.bias.normal = torch.distributions.Normal(l.bias.zero, l.bias.one)
    ops.enable_kernel_timing(True)
    try:
        fast(input_ids=ids, attention_mask=mask)
        torch.cuda.synchronize()
        ran = set(ops.kernel_timing_summary())
    finally:
        ops.enable_kernel_timing(False)
    assert "attention_fwd" not in ran
