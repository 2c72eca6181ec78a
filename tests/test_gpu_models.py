"""GPU parity tests at model scale (run with `-m gpu` on the B200 box): full-width BERT-base against the oracle's
sequential S-loop, the all-Bayesian (Linear + Embedding + LayerNorm) conversion at BERT-large width, the row-sparse
Embedding kernels, multi-tensor sampling over every Bayesian layer, reproducible resume, and the two-rank NCCL runs.

The oracle (oracle/bayes_oracle.py) is the CPU restatement of the reference pinned by tests/golden/*.npz; both sides
consume the SAME injected eps.  Tolerances: 1e-5 relative in fp32 mode (times a small depth factor for logits that
went through 12 encoder layers), 1e-2 in bf16 GEMM mode (north star).
"""
import copy
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_err
from oracle import bayes_oracle as O

import bayeformers_b200 as bf
import bayeformers_b200.nn as bnn
from bayeformers_b200 import ops
from bayeformers_b200.nn.layers.common import BayesianLayer

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FP32_TOL = 1e-5
BF16_TOL = 1e-2


class FixedEps:
    def __init__(self, queue):
        self.queue = [torch.as_tensor(q) for q in queue]

    def sample(self, size):
        e = self.queue.pop(0)
        assert tuple(e.shape) == tuple(size)
        return e


def _perturb_biases(model, seed):
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():  # HF zero-inits biases / unit LayerNorm weights: give MOPED non-degenerate values
        for n, p in model.named_parameters():
            if n.endswith("bias") or "LayerNorm.weight" in n:
                p.add_(torch.randn(p.shape, generator=gen) * 0.02)


def _inject_same_eps(bm, om, S, seed):
    """Same eps into our model (FixedEps on every Gaussian, S draws each) and into the oracle twin (per-module queues:
    sample-major, weight then bias -- the order a sequential S-loop consumes them)."""
    gen = torch.Generator().manual_seed(seed)
    omods = dict(om.named_modules())
    n_layers = 0
    for name, mod in bm.model.named_modules():
        if not isinstance(mod, BayesianLayer):
            continue
        n_layers += 1
        ew = [torch.randn(mod.weight.mu.shape, generator=gen) for _ in range(S)]
        has_b = isinstance(getattr(mod, "bias", None), bnn.Gaussian)
        eb = [torch.randn(mod.bias.mu.shape, generator=gen) for _ in range(S)] if has_b else None
        mod.weight.normal = FixedEps(ew)
        if has_b:
            mod.bias.normal = FixedEps(eb)
        queue = []
        for s in range(S):
            queue.append(ew[s])
            if has_b:
                queue.append(eb[s])
        omods[name].eps = O.EpsSource(preset=queue)
    return n_layers


def _oracle_loop(om, call, S):
    outs, lps, lqs = [], [], []
    for _ in range(S):
        outs.append(call(om))
        lps.append(torch.as_tensor(O.model_log_prior(om)))
        lqs.append(torch.as_tensor(O.model_log_variational_posterior(om)))
    return torch.stack(outs), torch.stack(lps), torch.stack(lqs)


# ------------------------------------------------------------------ full-width BERT-base (BASELINE configs[2] shape)
@pytest.mark.parametrize("mode", ["fp32", "fp32x3", "bf16", "bf16_bench"])
def test_bert_base_full_width_matches_oracle_s_loop(mode):
    """BERT-base (12 layers, H=768, 12 heads, FF=3072), T=128, S=4, B=2, to_bayesian(delta=0.05, freeze=True):
    per-sample logits and log-probs of ONE folded forward against the oracle's sequential S-loop on the CPU
    (pattern of examples/bert_glue.py:56-73), plus rho gradients of the first / last layers in fp32 mode.
    "bf16" = the north star's bf16 GEMM mode (bf16 operands, fp32 accumulate, fp32 host model): 1e-2.  "bf16_bench" =
    the configuration bench.py times: additionally the host model's embeddings / LayerNorm outputs and every activation
    between the layers travel in bf16 (cast_frequentist_), with the fused GELU, fused dropout+residual+LayerNorm (eval
    mode here), native LayerNorm and gradient sinks on -- twelve layers of bf16 activations sit at ~1.2e-2, checked
    against 2e-2 and reported."""
    from transformers import BertConfig, BertForSequenceClassification
    torch.manual_seed(0)
    cfg = BertConfig(num_labels=2)
    model = BertForSequenceClassification(cfg).eval()
    _perturb_biases(model, 1)
    S, B, Tn, n_batches = 4, 2, 128, 1000
    gen = torch.Generator().manual_seed(2)
    ids = torch.randint(0, cfg.vocab_size, (B, Tn), generator=gen)
    labels = torch.randint(0, 2, (B,), generator=gen)
    bench_cfg = mode == "bf16_bench"
    if bench_cfg:
        mode = "bf16"
    bm = bf.to_bayesian(model, delta=0.05, freeze=True, gemm_dtype=mode)
    if bench_cfg:
        bf.accelerate_host_(bm, layernorm=True, fuse_gelu=True, fuse_residual=True, grad_sinks=True)
    bm = bm.eval().to(DEV)
    if bench_cfg:
        bf.cast_frequentist_(bm, torch.bfloat16)
    om = O.oracle_convert(model, 0.05, True).eval()
    assert _inject_same_eps(bm, om, S, 3) == 74
    try:
        with bf.mc_samples(S):
            logits = bm(input_ids=ids.to(DEV).repeat(S, 1)).logits
        raw = logits.float().view(S, B, -1)
        lp_s, lq_s = bm.log_prior(), bm.log_variational_posterior()
        assert lp_s.shape == (S,) and lq_s.shape == (S,)
        nll = torch.nn.functional.cross_entropy(raw.mean(0), labels.to(DEV))
        loss = (lq_s.mean() - lp_s.mean()) / n_batches + nll
        loss.backward()
    finally:
        bf.runtime.enable_grad_sinks(False)
    o_raw, o_lp, o_lq = _oracle_loop(om, lambda m: m(input_ids=ids).logits, S)
    o_nll = torch.nn.functional.cross_entropy(o_raw.mean(0), labels)
    o_loss = (o_lq.mean() - o_lp.mean()) / n_batches + o_nll
    o_loss.backward()
    tol = (2 * BF16_TOL if bench_cfg else BF16_TOL) if mode == "bf16" else FP32_TOL
    e_logits = rel_err(raw.detach().cpu().numpy(), o_raw.detach().numpy())
    e_lp, e_lq = rel_err(lp_s.cpu().numpy(), o_lp.numpy()), rel_err(lq_s.cpu().numpy(), o_lq.numpy())
    tag = "bf16_bench" if bench_cfg else mode
    print(f"[bert-base {tag}] logits {e_logits:.2e}  log p {e_lp:.2e}  log q {e_lq:.2e}")
    assert e_logits < (tol if mode == "bf16" else 5 * tol)  # 12 encoder layers deep
    assert e_lp < FP32_TOL and e_lq < FP32_TOL                # the log-probs never depend on the GEMM dtype
    assert abs(float(loss) - float(o_loss)) <= (1e-3 if mode == "bf16" else 1e-5) * abs(float(o_loss))
    ours = {n: m for n, m in bm.model.named_modules() if isinstance(m, bnn.Linear)}
    theirs = {n: m for n, m in om.named_modules() if isinstance(m, O.OracleLinear)}
    gtol = 5e-2 if mode == "bf16" else 5e-4  # gradients went back through 12 layers of attention / LayerNorm
    for name in ("classifier", "bert.pooler.dense", "bert.encoder.layer.11.output.dense",
                 "bert.encoder.layer.11.attention.self.value", "bert.encoder.layer.0.intermediate.dense",
                 "bert.encoder.layer.0.attention.self.query"):
        e = rel_err(ours[name].weight.rho.grad.cpu().numpy(), theirs[name].w_rho.grad.numpy())
        print(f"[bert-base {tag}] d rho {name}: {e:.2e}")
        assert e < gtol, name
        assert ours[name].weight.mu.grad is None


# ------------------------------------------------------------------ all-Bayesian conversion at BERT-large width
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_all_bayesian_bert_large_width_matches_composed_oracle(mode):
    """BASELINE configs[4] in miniature: 2 encoder layers at BERT-large width (H=1024, 16 heads, FF=4096),
    to_bayesian(layers=TORCH2BAYE_ALL) -- every Linear, Embedding and LayerNorm Bayesian -- S=2 folded, against the
    composed oracle (the reference's Gaussian arithmetic + F.embedding / F.layer_norm)."""
    from transformers import BertConfig, BertForSequenceClassification
    torch.manual_seed(0)
    cfg = BertConfig(vocab_size=512, hidden_size=1024, num_hidden_layers=2, num_attention_heads=16,
                     intermediate_size=4096, max_position_embeddings=64, num_labels=2)
    model = BertForSequenceClassification(cfg).eval()
    _perturb_biases(model, 5)
    S, B, Tn = 2, 3, 32
    gen = torch.Generator().manual_seed(6)
    ids = torch.randint(0, cfg.vocab_size, (B, Tn), generator=gen)
    ids[0, :4] = 0  # padding_idx rows + repeated ids
    bm = bf.to_bayesian(model, delta=0.05, freeze=True, layers=bnn.TORCH2BAYE_ALL, gemm_dtype=mode)
    if mode == "bf16":
        bf.accelerate_host_(bm, layernorm=True, fuse_gelu=True, fuse_residual=True)
    bm = bm.eval().to(DEV)
    om = O.oracle_convert(model, 0.05, True, all_layers=True).eval()
    n_bayes = _inject_same_eps(bm, om, S, 7)
    assert n_bayes == 14 + 3 + 5  # (6 Linear x 2 layers + pooler + classifier) + 3 Embedding + 5 LayerNorm
    with bf.mc_samples(S):
        logits = bm(input_ids=ids.to(DEV).repeat(S, 1)).logits
    raw = logits.float().view(S, B, -1)
    lp_s, lq_s = bm.log_prior(), bm.log_variational_posterior()
    raw.square().sum().backward()
    o_raw, o_lp, o_lq = _oracle_loop(om, lambda m: m(input_ids=ids).logits, S)
    o_raw.square().sum().backward()
    tol = FP32_TOL if mode == "fp32" else BF16_TOL
    e = rel_err(raw.detach().cpu().numpy(), o_raw.detach().numpy())
    print(f"[all-bayes {mode}] logits {e:.2e}")
    assert e < 3 * tol
    assert rel_err(lp_s.cpu().numpy(), o_lp.numpy()) < FP32_TOL
    assert rel_err(lq_s.cpu().numpy(), o_lq.numpy()) < FP32_TOL
    omods = dict(om.named_modules())
    gtol = 2e-4 if mode == "fp32" else 5e-2
    for name, mod in bm.model.named_modules():
        if isinstance(mod, (bnn.Embedding, bnn.LayerNorm)) or name in ("classifier", "bert.encoder.layer.0.output.dense"):
            want = omods[name].w_rho.grad
            e = rel_err(mod.weight.rho.grad.cpu().numpy(), want.numpy())
            print(f"[all-bayes {mode}] d rho {name}: {e:.2e}")
            assert e < gtol, name


# ------------------------------------------------------------------ Embedding kernels (row A9)
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("kl", [False, True])
def test_embedding_kernels_against_materialised_table(out_dtype, kl):
    """bf_embedding_fwd / _bwd with the Philox stream: the looked-up rows and the row-sparse gradients must equal what
    sampling the WHOLE table with the same (seed, step, tensor_id) and F.embedding give.  The id list has long runs of
    one row (they cross several 256-token chunks of the backward), rows that occur once, padding ids and unused rows."""
    V, H, S, Bn, Tn = 300, 64, 3, 8, 160
    torch.manual_seed(1)
    emb = torch.nn.Embedding(V, H, padding_idx=2)
    be = bnn.Embedding.from_frequentist(emb, delta=0.05, freeze=False).to(DEV)
    be.kl_grad = kl
    be.gemm_dtype = out_dtype
    bf.manual_seed(4242)
    gen = torch.Generator().manual_seed(3)
    ids = torch.randint(0, V, (Bn, Tn), generator=gen)
    ids[:, :100] = 7          # one row, 800 tokens per sample: spans > 3 chunks
    ids[0, 100:110] = 2       # padding_idx
    ids[1, 100:140] = 11
    ids = ids.repeat(S, 1).to(DEV)
    with bf.mc_samples(S):
        out = be(ids)
    assert out.dtype == out_dtype and out.shape == (S * Bn, Tn, H)
    gy = torch.randn(out.shape, generator=gen).to(DEV)
    c = 1e-3
    loss = (out.float() * gy).sum()
    if kl:
        loss = loss + c * (be.live_log_variational_posterior - be.live_log_prior).sum()
    loss.backward()
    stream, _ = be._last_streams
    eps = torch.stack([ops.philox_normal(V * H, stream.seed, stream.step, stream.tensor_id, s, DEV).view(V, H)
                       for s in range(S)])
    mu = be.weight.mu.detach().double().requires_grad_()
    rho = be.weight.rho.detach().double().requires_grad_()
    W = mu + torch.nn.functional.softplus(rho) * eps.double()        # [S, V, H] float64
    ref = torch.cat([torch.nn.functional.embedding(ids.view(S, Bn, Tn)[s], W[s], padding_idx=2) for s in range(S)])
    rloss = (ref * gy.double()).sum()
    if kl:
        sig = torch.nn.functional.softplus(rho)
        sp = float(torch.nn.functional.softplus(torch.tensor(1.0, dtype=torch.float64)))
        lq = (-0.5 * np.log(2 * np.pi) - sig.log() - (W - mu) ** 2 / (2 * sig ** 2)).sum((1, 2))
        lp = (-0.5 * np.log(2 * np.pi) - np.log(sp) - (W - emb.weight.detach().double().to(DEV)) ** 2 / (2 * sp ** 2)).sum((1, 2))
        rloss = rloss + c * (lq - lp).sum()
        assert rel_err(be.log_variational_posterior_samples.cpu().numpy(), lq.detach().cpu().numpy()) < FP32_TOL
        assert rel_err(be.log_prior_samples.cpu().numpy(), lp.detach().cpu().numpy()) < FP32_TOL
    rloss.backward()
    tol = 2e-6 if out_dtype == torch.float32 else 4e-3
    assert rel_err(out.detach().float().cpu().numpy(), ref.detach().cpu().numpy()) < tol
    # a bf16 output makes autograd round the incoming gradient to bf16 (2^-9 per element)
    gtol = 1e-5 if out_dtype == torch.float32 else 4e-3
    assert rel_err(be.weight.rho.grad.cpu().numpy(), rho.grad.cpu().numpy()) < gtol
    assert rel_err(be.weight.mu.grad.cpu().numpy(), mu.grad.cpu().numpy()) < gtol
    if not kl:
        assert float(be.weight.rho.grad[2].abs().max()) == 0.0   # padding row
        assert float(be.weight.rho.grad[299].abs().max()) == 0.0 or (ids == 299).any()
    # run-to-run bit stability (no float atomics)
    g1 = be.weight.rho.grad.clone()
    be.weight.rho.grad = be.weight.mu.grad = None
    be.weight.step -= 1  # replay the same draw
    with bf.mc_samples(S):
        out2 = be(ids)
    loss2 = (out2.float() * gy).sum()
    if kl:
        loss2 = loss2 + c * (be.live_log_variational_posterior - be.live_log_prior).sum()
    loss2.backward()
    assert torch.equal(out, out2) and torch.equal(g1, be.weight.rho.grad)


# ------------------------------------------------------------------ multi-tensor sampling over every Bayesian layer
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_presample_covers_embedding_and_layernorm(mode):
    """enable_presample on an all-Bayesian model: Linear, LayerNorm and Embedding tensors drawn / reduced by ONE
    bf_sample_kl_fwd_multi launch give the same outputs, log-probs and gradients as the per-layer kernels fed the same
    Philox streams."""
    from transformers import BertConfig, BertForSequenceClassification
    torch.manual_seed(3)
    cfg = BertConfig(vocab_size=200, hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                     intermediate_size=512, max_position_embeddings=32, num_labels=2)
    model = BertForSequenceClassification(cfg).eval()
    _perturb_biases(model, 4)
    S, B, Tn = 3, 4, 16
    ids = torch.randint(0, 200, (B, Tn), generator=torch.Generator().manual_seed(5)).to(DEV)
    bf.manual_seed(99)
    a = bf.to_bayesian(model, delta=0.05, freeze=True, layers=bnn.TORCH2BAYE_ALL, gemm_dtype=mode, kl_grad=True).eval().to(DEV)
    b = copy.deepcopy(a)
    bf.enable_presample(a)
    l0 = ops.stats["launches"]
    with bf.mc_samples(S):
        ya = a(input_ids=ids.repeat(S, 1)).logits
    (ya.float().square().sum() + 1e-3 * (a.log_variational_posterior() - a.log_prior()).sum()).backward()
    n_bayes = len(a.bayesian_children)
    assert n_bayes == 14 + 3 + 5
    # replay the same eps through the per-layer kernels of the twin
    for la_, lb_ in zip(a.bayesian_children, b.bayesian_children):
        ws, bs = la_._last_streams
        assert ws.step & 0x80000000, type(la_).__name__
        n = lb_.weight.mu.numel()
        lb_.weight.normal = FixedEps([ops.philox_normal(n, ws.seed, ws.step, ws.tensor_id, s, DEV).view_as(lb_.weight.mu)
                                      for s in range(S)])
        if isinstance(getattr(lb_, "bias", None), bnn.Gaussian):
            nb = lb_.bias.mu.numel()
            lb_.bias.normal = FixedEps([ops.philox_normal(nb, bs.seed, bs.step, bs.tensor_id, s, DEV) for s in range(S)])
    with bf.mc_samples(S):
        yb = b(input_ids=ids.repeat(S, 1)).logits
    (yb.float().square().sum() + 1e-3 * (b.log_variational_posterior() - b.log_prior()).sum()).backward()
    tol = FP32_TOL if mode == "fp32" else BF16_TOL
    assert rel_err(ya.detach().float().cpu().numpy(), yb.detach().float().cpu().numpy()) < tol
    assert rel_err(a.log_prior().detach().cpu().numpy(), b.log_prior().detach().cpu().numpy()) < 2e-6
    assert rel_err(a.log_variational_posterior().detach().cpu().numpy(), b.log_variational_posterior().detach().cpu().numpy()) < 2e-6
    for (na, pa), (nb_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert (pa.grad is None) == (pb.grad is None), na
        if pa.grad is not None:
            assert rel_err(pa.grad.cpu().numpy(), pb.grad.cpu().numpy()) < max(tol, 1e-4), na
    # the registered per-layer scalars are 0-dim after a folded forward: reference-shaped checkpoints
    sd = a.state_dict()
    assert all(v.dim() == 0 for k, v in sd.items() if k.endswith("log_prior") or k.endswith("log_variational_posterior"))


# ------------------------------------------------------------------ reproducible resume
def test_rng_state_resume_reproduces_the_next_step():
    """bf.rng_state / load_rng_state + ClipAdamW.state_dict: a run resumed from a checkpoint taken after step 1 draws
    the same eps and dropout masks and lands on bit-identical parameters after step 2."""
    from transformers import BertConfig, BertForSequenceClassification
    torch.manual_seed(8)
    cfg = BertConfig(vocab_size=100, hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                     intermediate_size=512, max_position_embeddings=32, num_labels=2)
    model = BertForSequenceClassification(cfg)
    _perturb_biases(model, 9)
    S, B, Tn = 2, 4, 16
    ids = torch.randint(0, 100, (B, Tn), generator=torch.Generator().manual_seed(1)).to(DEV)
    labels = torch.tensor([0, 1, 1, 0], device=DEV)

    def build():
        bf.manual_seed(2024)
        bm = bf.to_bayesian(model, delta=0.05, freeze=True, gemm_dtype="bf16", kl_grad=True)
        bf.accelerate_host_(bm, fuse_residual=True)
        bm = bm.to(DEV).train()
        bf.enable_presample(bm)
        bf.cast_frequentist_(bm, torch.bfloat16)
        opt = bf.optim.ClipAdamW([p for p in bm.parameters() if p.requires_grad], lr=1e-3, max_grad_norm=1.0)
        return bm, opt

    def step(bm, opt):
        bf.advance_step()
        opt.zero_grad()
        with bf.mc_samples(S):
            logits = bm(input_ids=ids.repeat(S, 1)).logits
        loss = torch.nn.functional.cross_entropy(logits.float().view(S, B, -1).mean(0), labels)
        loss = loss + (bm.log_variational_posterior().mean() - bm.log_prior().mean()) / 100
        loss.backward()
        opt.step()
        return loss.detach().clone()

    bf.enable_device_step(DEV)
    try:
        bm, opt = build()
        step(bm, opt)
        ckpt = {"model": copy.deepcopy(bm.state_dict()), "rng": bf.rng_state(bm), "optimizer": copy.deepcopy(opt.state_dict()),
                "torch_cuda_rng": torch.cuda.get_rng_state(DEV)}  # the host model's own dropout draws from torch's generator
        loss2 = step(bm, opt)
        want = {k: v.clone() for k, v in bm.state_dict().items()}
        # "new process": fresh objects (fresh stream ids, counters at zero), then restore
        bf.disable_device_step()
        bf.enable_device_step(DEV)
        bm2, opt2 = build()
        bm2.load_state_dict(ckpt["model"], strict=True)
        opt2.load_state_dict(ckpt["optimizer"])
        bf.load_rng_state(bm2, ckpt["rng"])
        torch.cuda.set_rng_state(ckpt["torch_cuda_rng"], DEV)
        loss2b = step(bm2, opt2)
        assert torch.equal(loss2, loss2b)
        for k, v in bm2.state_dict().items():
            assert torch.equal(v, want[k]), k
    finally:
        bf.disable_device_step()


# ------------------------------------------------------------------ two ranks over NCCL
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_two_rank_nccl_batch_and_sample_sharding():
    """Launches tests/multirank_worker.py on 2 ranks (torch.distributed.run, NCCL): identical sampled weights on both
    ranks, all-reduced rho gradients == mean of the single-rank gradients (batch sharding), and the sample-sharded
    step == the single-process S-sample step.  The worker asserts; this test checks its exit status."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "tests", "multirank_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0
    assert "MULTIRANK OK" in r.stdout


# ------------------------------------------------------------------ no per-step memory growth
def test_eager_training_steps_do_not_leak_device_memory():
    """Eager training steps with every extension on (multi-tensor sampling, fused blocks, gradient sinks, GELU links,
    kl_grad): the device memory in use after a step must not grow from step to step WITHOUT help from the garbage
    collector.  (An autograd node whose ctx reaches one of its own outputs is a reference cycle; one such cycle used to
    pin a whole sampled-weight arena and the GELU pre-activations of every step.)"""
    import gc
    from transformers import BertConfig, BertForSequenceClassification
    torch.manual_seed(0)
    cfg = BertConfig(vocab_size=100, hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                     intermediate_size=1024, max_position_embeddings=32, num_labels=2)
    model = BertForSequenceClassification(cfg)
    bm = bf.to_bayesian(model, delta=0.05, freeze=True, layers=bnn.TORCH2BAYE_ALL, gemm_dtype="bf16", kl_grad=True)
    bf.accelerate_host_(bm, fuse_residual=True, grad_sinks=True)
    bm = bm.to(DEV).train()
    bf.enable_presample(bm)
    bf.cast_frequentist_(bm, torch.bfloat16)
    opt = bf.optim.ClipAdamW([p for p in bm.parameters() if p.requires_grad], lr=1e-4, max_grad_norm=1.0)
    S, B, Tn = 4, 64, 32  # 8192 folded rows: enough 256 x 256 tiles for the fused GELU / GELU' kernels
    ids = torch.randint(0, 100, (B, Tn), device=DEV)
    labels = torch.randint(0, 2, (B,), device=DEV)
    used = []
    gc.collect()
    gc.disable()
    try:
        for _ in range(5):
            opt.zero_grad()
            with bf.mc_samples(S):
                logits = bm(input_ids=ids.repeat(S, 1)).logits
            loss = torch.nn.functional.cross_entropy(logits.float().view(S, B, -1).mean(0), labels)
            loss = loss + (bm.log_variational_posterior().mean() - bm.log_prior().mean()) / 100
            loss.backward()
            opt.step()
            del logits, loss
            torch.cuda.synchronize()
            used.append(torch.cuda.memory_allocated())
    finally:
        gc.enable()
        bf.runtime.enable_grad_sinks(False)
    assert used[4] == used[3] == used[2], used


# ------------------------------------------------------------------ sigma cache kept by the optimizer
def test_sigma_cache_written_by_the_optimizer_changes_no_bit():
    """ClipAdamW(model=...) writes softplus(updated rho) next to its update and the multi-tensor sampling kernel reads
    that cache instead of rho: draws, log-probs and the next update must be bit-identical to the uncached path, before
    and after optimizer steps, and a torch-side change of rho must invalidate the cache."""
    torch.manual_seed(4)
    net = torch.nn.Sequential(torch.nn.Linear(256, 512), torch.nn.Tanh(), torch.nn.Linear(512, 64))
    x = torch.randn(3 * 16, 256, device=DEV)
    S = 3
    models, opts = [], []
    for cached in (False, True):
        bf.manual_seed(31)
        bm = bf.to_bayesian(copy.deepcopy(net), delta=0.05, freeze=True, gemm_dtype="bf16", kl_grad=True).to(DEV)
        bf.enable_presample(bm)
        opts.append(bf.optim.ClipAdamW([p for p in bm.parameters() if p.requires_grad], lr=1e-2, max_grad_norm=1.0,
                                       model=bm if cached else None))
        models.append(bm)
    bf.load_rng_state(models[1], bf.rng_state(models[0]))  # same stream ids and counters: both models draw the same eps
    gauss = [g for g in models[1].modules() if isinstance(g, bnn.Gaussian) and g.rho.requires_grad and g.rho.grad is None]
    assert all(g.sigma_cache() is not None for g in gauss if any(g.rho is p for p in opts[1].params))
    for step in range(3):
        outs = []
        for bm, opt in zip(models, opts):
            opt.zero_grad()
            with bf.mc_samples(S):
                y = bm(x)
            loss = y.float().square().mean() + 1e-3 * (bm.log_variational_posterior() - bm.log_prior()).mean()
            loss.backward()
            opt.step()
            outs.append((y.detach().clone(), bm.log_prior().detach().clone(), bm.log_variational_posterior().detach().clone()))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2]), step
        for pa, pb in zip(models[0].parameters(), models[1].parameters()):
            assert torch.equal(pa, pb)
    lin = [m for m in models[1].modules() if isinstance(m, bnn.Linear)][0]
    sig = lin.weight.sigma_cache()
    assert sig is not None and torch.equal(sig, bnn.Gaussian.sigma.fget(lin.weight).detach()) or \
        rel_err(sig.cpu().numpy(), torch.nn.functional.softplus(lin.weight.rho.detach()).cpu().numpy()) < 1e-6
    with torch.no_grad():
        lin.weight.rho.add_(0.01)  # torch-side change: the cache must not be used any more
    assert lin.weight.sigma_cache() is None
    with bf.mc_samples(S):
        models[1](x)  # uncached kernels for this tensor: no error, values follow the new rho
    opts[1].zero_grad()
    with bf.mc_samples(S):
        models[1](x).float().sum().backward()
    opts[1].step()
    assert lin.weight.sigma_cache() is not None  # the optimizer's write made it current again
