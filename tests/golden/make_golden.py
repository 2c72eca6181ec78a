"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, CPU):

    python tests/golden/make_golden.py

The reference (yliess86/BayeFormers) is imported from /root/reference; eps is
injected without touching its code by replacing the `normal` attribute of each
`Gaussian` with a stub that returns preset tensors (hooks
bayeformers/nn/parameters/gaussian.py:100).  Everything written here is a
small .npz of inputs + reference outputs; the reference's sources are never
copied.  The GPU box has no /root/reference: tests read only these files.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
os.environ.setdefault("HF_HUB_OFFLINE", "1")

import bayeformers  # noqa: E402
import bayeformers.nn as rbnn  # noqa: E402
from bayeformers import to_bayesian as ref_to_bayesian  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


class FixedEps:
    """Stand-in for torch.distributions.Normal: pops preset tensors."""

    def __init__(self, queue):
        self.queue = list(queue)

    def sample(self, size):
        e = self.queue.pop(0)
        assert tuple(e.shape) == tuple(size)
        return e


def npy(t):
    return t.detach().cpu().numpy().copy()


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path)} bytes)")


# --------------------------------------------------------------------------- #
def gen_gaussian_kat():
    """Gaussian.sample / .log_prob on the SURVEY 8c hand-picked values plus a
    seeded random case covering the softplus threshold and tiny sigmas."""
    out = {}
    g = rbnn.Gaussian(torch.Size((3,)))
    g.mu.data = torch.tensor([0.0, 1.0, -2.0])
    g.rho.data = torch.tensor([0.0, -5.0, 25.0])
    eps = torch.tensor([1.0, -1.0, 0.5])
    g.normal = FixedEps([eps])
    w = g.sample()
    out.update(kat_mu=npy(g.mu), kat_rho=npy(g.rho), kat_eps=npy(eps), kat_sigma=npy(g.sigma),
               kat_w=npy(w), kat_logq=npy(g.log_prob(w)))

    gen = torch.Generator().manual_seed(1234)
    n = 4099  # deliberately not a multiple of 4
    mu = torch.empty(n).uniform_(-0.2, 0.2, generator=gen)
    rho = torch.empty(n).uniform_(-8.0, 3.0, generator=gen)
    rho[:8] = torch.tensor([19.5, 20.0, 20.5, 30.0, -30.0, -40.0, 0.0, 1.0])
    eps = torch.randn(n, generator=gen)
    g = rbnn.Gaussian(torch.Size((n,)))
    g.mu.data, g.rho.data = mu, rho
    g.normal = FixedEps([eps])
    w = g.sample()
    out.update(rnd_mu=npy(mu), rnd_rho=npy(rho), rnd_eps=npy(eps), rnd_sigma=npy(g.sigma),
               rnd_w=npy(w), rnd_logq=npy(g.log_prob(w)))
    save("gaussian.npz", **out)


def gen_mixture_kat():
    prior = rbnn.DEFAULT_SCALED_GAUSSIAN_MIXTURE
    w = torch.tensor([0.0, 0.001, 0.01, 0.05, 1.0, -3.0, 14.0])
    # per-element values via 1-element calls (log_prob sums)
    per = torch.stack([prior.log_prob(w[i:i + 1]) for i in range(len(w))])
    gen = torch.Generator().manual_seed(99)
    wr = torch.randn(5000, generator=gen) * 0.3
    wr[:6] = torch.tensor([0.0, 0.03, -0.033, 0.04, 5.0, -9.0])
    custom = rbnn.ScaledGaussianMixture(0.25, 0.7, 0.05)
    save("mixture.npz", kat_w=npy(w), kat_logp_elem=npy(per),
         pi=np.float32(prior.pi.item()), sigma1=np.float32(prior.sigma1.item()),
         sigma2=np.float32(prior.sigma2.item()),
         rnd_w=npy(wr), rnd_logp=npy(prior.log_prob(wr)),
         custom_params=np.array([0.25, 0.7, 0.05], dtype=np.float32), custom_logp=npy(custom.log_prob(wr)))


def _linear_case(tag, in_f, out_f, batch, delta, freeze, bias, seed, out):
    torch.manual_seed(seed)
    lin = torch.nn.Linear(in_f, out_f, bias=bias)
    with torch.no_grad():
        lin.weight.mul_(3.0)  # a wider spread than kaiming-uniform alone
        if delta is not None:
            lin.weight[0, 0] = 0.0  # exercises the -inf -> 0 rule
            lin.weight[0, 1] = 1e-7
    w0 = lin.weight.detach().clone()
    b0 = lin.bias.detach().clone() if bias else None
    layer = rbnn.Linear.from_frequentist(lin, delta=delta, freeze=freeze)
    x = torch.randn(batch, in_f, requires_grad=True)
    eps_w = torch.randn(out_f, in_f)
    eps_b = torch.randn(out_f) if bias else None
    layer.weight.normal = FixedEps([eps_w])
    if bias:
        layer.bias.normal = FixedEps([eps_b])
    y = layer(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    out[f"{tag}_w0"] = npy(w0)
    if bias:
        out[f"{tag}_b0"] = npy(b0)
        out[f"{tag}_eps_b"] = npy(eps_b)
        out[f"{tag}_b_mu"] = npy(layer.bias.mu)
        out[f"{tag}_b_rho"] = npy(layer.bias.rho)
        out[f"{tag}_g_b_rho"] = npy(layer.bias.rho.grad)
        if layer.bias.mu.grad is not None:
            out[f"{tag}_g_b_mu"] = npy(layer.bias.mu.grad)
    out[f"{tag}_x"] = npy(x)
    out[f"{tag}_eps_w"] = npy(eps_w)
    out[f"{tag}_gy"] = npy(gy)
    out[f"{tag}_w_mu"] = npy(layer.weight.mu)
    out[f"{tag}_w_rho"] = npy(layer.weight.rho)
    out[f"{tag}_y"] = npy(y)
    out[f"{tag}_log_prior"] = npy(layer.log_prior)
    out[f"{tag}_log_q"] = npy(layer.log_variational_posterior)
    out[f"{tag}_g_x"] = npy(x.grad)
    out[f"{tag}_g_w_rho"] = npy(layer.weight.rho.grad)
    if layer.weight.mu.grad is not None:
        out[f"{tag}_g_w_mu"] = npy(layer.weight.mu.grad)
    out[f"{tag}_meta"] = np.array([in_f, out_f, batch, -1.0 if delta is None else delta,
                                   float(freeze), float(bias)], dtype=np.float64)
    if delta is not None:
        out[f"{tag}_prior_mu"] = npy(layer.weight_prior.mu)
        out[f"{tag}_prior_rho"] = npy(layer.weight_prior.rho)
        out[f"{tag}_mu_requires_grad"] = np.array([layer.weight.mu.requires_grad])


def gen_linear():
    """bnn.Linear forward (y, log_prior, log_q) and backward (dX, dmu, drho)
    with injected eps: default init + mixture prior, MOPED trainable-mu,
    MOPED frozen-mu, and a bias-less layer."""
    out = {}
    # delta=None: from_frequentist discards the weights and keeps the uniform init
    _linear_case("default", 40, 24, 9, None, False, True, 11, out)
    _linear_case("moped", 48, 20, 7, 0.05, False, True, 12, out)
    _linear_case("moped_frozen", 36, 28, 5, 0.05, True, True, 13, out)
    _linear_case("nobias", 32, 16, 6, None, False, False, 14, out)
    # MOPED (Gaussian prior) with in / out features that are multiples of 8, i.e. shapes the TMA-fed tcgen05 kernels
    # take: trainable mu, frozen mu, and a frozen-mu case with several k-steps, ragged tiles and a ragged row count
    _linear_case("moped_tc", 48, 24, 7, 0.05, False, True, 15, out)
    _linear_case("moped_frozen_tc", 40, 32, 5, 0.05, True, True, 16, out)
    _linear_case("moped_frozen_tiles", 136, 72, 130, 0.05, True, True, 17, out)
    save("linear.npz", **out)


def gen_moped():
    """MOPED rho for special values (bit patterns are what matters)."""
    vals = torch.tensor([0.0, 1e-7, 1e-6, 1e-3, 0.02, -0.02, 0.5, -1.5, 10.0, 2000.0, -0.0, 3e-8], dtype=torch.float32)
    gen = torch.Generator().manual_seed(7)
    rnd = torch.randn(2048, generator=gen) * 0.02
    w = torch.cat([vals, rnd]).reshape(-1, 4)
    out = {"w": npy(w)}
    for delta in (0.05, 0.1, 0.01):
        lin = torch.nn.Linear(4, w.shape[0], bias=True)
        lin.weight.data = w.clone()
        lin.bias.data = w[:, 0].clone()
        layer = rbnn.Linear.from_frequentist(lin, delta=delta, freeze=True)
        out[f"rho_w_{delta}"] = npy(layer.weight.rho)
        out[f"rho_b_{delta}"] = npy(layer.bias.rho)
        out[f"mu_w_{delta}"] = npy(layer.weight.mu)
    save("moped.npz", **out)


class TinyMLP(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.body = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.ReLU(), torch.nn.Linear(16, 8, bias=False))
        self.head = torch.nn.Linear(8, 3)

    def forward(self, x):
        return self.head(torch.relu(self.body(x)))


def gen_to_bayesian():
    """to_bayesian on a small nested model: state_dict keys, shapes,
    requires_grad flags, values, for (delta=None) and (delta, freeze)."""
    out = {}
    for tag, kw in (("plain", {}), ("moped", {"delta": 0.05, "freeze": True}), ("moped_unfrozen", {"delta": 0.1})):
        torch.manual_seed(21)
        m = TinyMLP()
        torch.manual_seed(22)  # controls the uniform draws inside the conversion
        bm = ref_to_bayesian(m, **kw)
        sd = bm.state_dict()
        out[f"{tag}_keys"] = np.array(list(sd.keys()))
        out[f"{tag}_freq_keys"] = np.array(list(m.state_dict().keys()))
        for k, v in m.state_dict().items():
            out[f"{tag}_freq::{k}"] = npy(v)
        for k, v in sd.items():
            out[f"{tag}::{k}"] = npy(v)
        out[f"{tag}_requires_grad"] = np.array([f"{n}={int(p.requires_grad)}" for n, p in bm.named_parameters()])
        out[f"{tag}_rng_after"] = torch.get_rng_state().numpy()[:64].copy()
        # one forward with injected eps through the whole model
        gen = torch.Generator().manual_seed(5)
        x = torch.randn(4, 12, generator=gen)
        eps_list = []
        for mod in bm.modules():
            if isinstance(mod, rbnn.Linear):
                ew = torch.randn(mod.weight.mu.shape, generator=gen)
                mod.weight.normal = FixedEps([ew])
                eps_list.append(ew)
                if isinstance(mod.bias, rbnn.Gaussian):
                    eb = torch.randn(mod.bias.mu.shape, generator=gen)
                    mod.bias.normal = FixedEps([eb])
                    eps_list.append(eb)
        y = bm(x)
        out[f"{tag}_x"] = npy(x)
        out[f"{tag}_y"] = npy(y)
        out[f"{tag}_log_prior"] = npy(bm.log_prior())
        out[f"{tag}_log_q"] = npy(bm.log_variational_posterior())
        for i, e in enumerate(eps_list):
            out[f"{tag}_eps{i}"] = npy(e)
    save("to_bayesian.npz", **out)


def gen_kl_grad():
    """kl_grad=True oracle: autograd through the reference's own log_prob
    classes (no `.data` detach), for the mixture and the Gaussian prior."""
    gen = torch.Generator().manual_seed(3)
    n = 1537
    mu = (torch.randn(n, generator=gen) * 0.05).requires_grad_()
    rho = torch.empty(n).uniform_(-6.0, -3.0, generator=gen).requires_grad_()
    eps = torch.randn(n, generator=gen)
    out = {"mu": npy(mu), "rho": npy(rho), "eps": npy(eps)}
    for tag in ("mixture", "gaussian"):
        g = rbnn.Gaussian(torch.Size((n,)))
        g.mu.data, g.rho.data = mu.detach().clone(), rho.detach().clone()
        g.normal = FixedEps([eps])
        if tag == "mixture":
            prior = rbnn.DEFAULT_SCALED_GAUSSIAN_MIXTURE
        else:
            prior = rbnn.Gaussian(torch.Size((n,)))
            prior.mu.data = (mu.detach() + 0.01).clone()
            prior.rho.data = torch.ones(n)
            out["gaussian_prior_mu"] = npy(prior.mu)
            out["gaussian_prior_rho"] = npy(prior.rho)
        w = g.sample()
        lq, lp = g.log_prob(w), prior.log_prob(w)
        c = 1.0 / 250.0
        (c * (lq - lp)).backward()
        out[f"{tag}_logq"], out[f"{tag}_logp"] = npy(lq), npy(lp)
        out[f"{tag}_g_mu"], out[f"{tag}_g_rho"] = npy(g.mu.grad), npy(g.rho.grad)
        out["kl_weight"] = np.float64(c)
    save("kl_grad.npz", **out)


def gen_tiny_bert():
    """2-layer toy BERT through the reference S-loop (S=3) with injected eps:
    per-sample logits, per-sample log_prior / log_q, loss and rho gradients of
    the pattern of examples/bert_glue.py:56-73,231-239."""
    from transformers import BertConfig, BertForSequenceClassification

    cfg = BertConfig(vocab_size=97, hidden_size=32, num_hidden_layers=2, num_attention_heads=2,
                     intermediate_size=64, max_position_embeddings=24, num_labels=2)
    torch.manual_seed(0)
    model = BertForSequenceClassification(cfg).eval()
    with torch.no_grad():  # HF zero-inits biases; perturb so MOPED sees non-trivial values
        for n_, p in model.named_parameters():
            if n_.endswith("bias"):
                p.add_(torch.randn_like(p) * 0.02)
    torch.manual_seed(1)
    bm = ref_to_bayesian(model, delta=0.05, freeze=True).eval()
    S, B, T, n_batches = 3, 4, 16, 100
    gen = torch.Generator().manual_seed(2)
    ids = torch.randint(0, cfg.vocab_size, (B, T), generator=gen)
    labels = torch.randint(0, 2, (B,), generator=gen)
    layers = [m for m in bm.modules() if isinstance(m, rbnn.Linear)]
    out = {"ids": npy(ids), "labels": npy(labels), "S": np.int64(S), "n_batches": np.int64(n_batches),
           "cfg": np.array([cfg.vocab_size, cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads,
                            cfg.intermediate_size, cfg.max_position_embeddings, cfg.num_labels])}
    for k, v in model.state_dict().items():
        out[f"freq::{k}"] = npy(v)
    eps_w = [[torch.randn(l.weight.mu.shape, generator=gen) for _ in range(S)] for l in layers]
    eps_b = [[torch.randn(l.bias.mu.shape, generator=gen) for _ in range(S)] for l in layers]
    for i, l in enumerate(layers):
        l.weight.normal = FixedEps(eps_w[i])
        l.bias.normal = FixedEps(eps_b[i])
        out[f"eps_w{i}"] = npy(torch.stack(eps_w[i]))
        out[f"eps_b{i}"] = npy(torch.stack(eps_b[i]))
    logits, lps, lqs = [], [], []
    for s in range(S):
        logits.append(bm(input_ids=ids).logits)
        lps.append(bm.log_prior())
        lqs.append(bm.log_variational_posterior())
    raw = torch.stack(logits)
    lp, lq = torch.stack(lps).mean(), torch.stack(lqs).mean()
    nll = torch.nn.functional.cross_entropy(raw.mean(0), labels)
    loss = (lq - lp) / n_batches + nll
    loss.backward()
    out.update(logits=npy(raw), log_prior=npy(torch.stack(lps)), log_q=npy(torch.stack(lqs)), loss=npy(loss))
    names = [n for n, m in bm.named_modules() if isinstance(m, rbnn.Linear)]
    out["layer_names"] = np.array(names)
    for i, l in enumerate(layers):
        out[f"g_w_rho{i}"] = npy(l.weight.rho.grad)
        out[f"g_b_rho{i}"] = npy(l.bias.rho.grad)
    save("tiny_bert.npz", **out)


if __name__ == "__main__":
    torch.set_num_threads(4)
    only = sys.argv[1:]
    if only:  # e.g. `python tests/golden/make_golden.py gen_linear` regenerates one file
        for name in only:
            globals()[name]()
        sys.exit(0)
    gen_gaussian_kat()
    gen_mixture_kat()
    gen_linear()
    gen_moped()
    gen_to_bayesian()
    gen_kl_grad()
    gen_tiny_bert()
