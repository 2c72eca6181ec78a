"""CPU: the host-side mirror of the reference API (conversion, MOPED, state_dict
names, RNG consumption) against the golden fixtures, and the C-ABI boundary
(library loads, exports every symbol include/*.h declares).  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden

import bayeformers_b200 as bf
import bayeformers_b200.nn as bnn
from bayeformers_b200 import _lib


class TinyMLP(torch.nn.Module):  # same architecture as tests/golden/make_golden.py
    def __init__(self):
        super().__init__()
        self.body = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.ReLU(), torch.nn.Linear(16, 8, bias=False))
        self.head = torch.nn.Linear(8, 3)

    def forward(self, x):
        return self.head(torch.relu(self.body(x)))


@pytest.mark.parametrize("tag,kw", [("plain", {}), ("moped", {"delta": 0.05, "freeze": True}),
                                    ("moped_unfrozen", {"delta": 0.1})])
def test_to_bayesian_matches_reference(tag, kw):
    g = load_golden("to_bayesian.npz")
    torch.manual_seed(21)
    m = TinyMLP()
    for k in g[f"{tag}_freq_keys"]:
        assert torch.equal(m.state_dict()[str(k)], torch.from_numpy(g[f"{tag}_freq::{k}"]))
    torch.manual_seed(22)
    bm = bf.to_bayesian(m, **kw)
    sd = bm.state_dict()
    assert list(sd.keys()) == [str(k) for k in g[f"{tag}_keys"]]          # names AND order
    for k, v in sd.items():
        ref = torch.from_numpy(g[f"{tag}::{k}"])
        assert v.shape == ref.shape, k
        assert torch.equal(v, ref), k                                        # bit-exact mu / rho / priors
    flags = [f"{n}={int(p.requires_grad)}" for n, p in bm.named_parameters()]
    assert flags == [str(s) for s in g[f"{tag}_requires_grad"]]
    # same consumption of torch's global generator as the reference (quirk Q12)
    assert np.array_equal(torch.get_rng_state().numpy()[:64], g[f"{tag}_rng_after"])
    assert isinstance(bm, bnn.Model) and len(bm.bayesian_children) == 3


def test_to_bayesian_is_not_in_place_and_exact_class_only():
    class MyLinear(torch.nn.Linear):
        pass

    net = torch.nn.Sequential(torch.nn.Linear(4, 4), MyLinear(4, 4))
    bm = bf.to_bayesian(net, delta=0.05)
    assert isinstance(net[0], torch.nn.Linear) and not isinstance(net[0], bnn.Linear)
    assert isinstance(bm.model[0], bnn.Linear)
    assert isinstance(bm.model[1], MyLinear)  # subclasses are skipped (quirk Q7)
    # root module is never replaced
    root = bf.to_bayesian(torch.nn.Linear(3, 3))
    assert isinstance(root.model, torch.nn.Linear)
    with pytest.warns(UserWarning):
        root.log_prior()


def test_moped_rule_and_aliasing():
    g = load_golden("moped.npz")
    w = torch.from_numpy(g["w"])
    for delta in (0.05, 0.1, 0.01):
        lin = torch.nn.Linear(4, w.shape[0])
        lin.weight.data = w.clone()
        lin.bias.data = w[:, 0].clone()
        layer = bnn.Linear.from_frequentist(lin, delta=delta, freeze=True)
        assert np.array_equal(layer.weight.rho.detach().numpy().view(np.uint32), g[f"rho_w_{delta}"].view(np.uint32))
        assert np.array_equal(layer.bias.rho.detach().numpy().view(np.uint32), g[f"rho_b_{delta}"].view(np.uint32))
        assert torch.equal(layer.weight.mu.data, w)
        # posterior mu, prior mu and the source weight share storage (quirk Q5)
        assert layer.weight.mu.data_ptr() == lin.weight.data_ptr() == layer.weight_prior.mu.data_ptr()
        assert torch.all(layer.weight_prior.rho == 1)
        assert not layer.weight.mu.requires_grad and layer.weight.rho.requires_grad


def test_delta_none_discards_pretrained_weights():
    lin = torch.nn.Linear(6, 5)
    torch.manual_seed(3)
    layer = bnn.Linear.from_frequentist(lin)
    assert layer.weight_prior is bnn.DEFAULT_SCALED_GAUSSIAN_MIXTURE
    assert float(layer.weight.mu.abs().max()) <= 0.2 and float(layer.weight.rho.max()) <= -4
    assert not torch.equal(layer.weight.mu.data, lin.weight.data)


def test_api_surface_matches_reference_exports():
    for name in ["Linear", "Model", "NoneParameter", "Parameter", "DEFAULT_SCALED_GAUSSIAN_MIXTURE", "Gaussian",
                 "ScaledGaussianMixture", "DEFAULT_UNIFORM", "Initialization", "Uniform", "TORCH2BAYE",
                 "Embedding", "LayerNorm"]:
        assert hasattr(bnn, name), name
    assert bnn.TORCH2BAYE == {torch.nn.Linear: bnn.Linear}
    p = bnn.DEFAULT_SCALED_GAUSSIAN_MIXTURE
    assert float(p.pi) == 0.5 and float(p.sigma1) == 1.0 and float(p.sigma2) == np.float32(np.exp(-6))
    assert bnn.DEFAULT_UNIFORM.mu_range == (-0.2, 0.2) and bnn.DEFAULT_UNIFORM.rho_range == (-5, -4)
    assert bnn.NoneParameter().sample() is None and bnn.NoneParameter().log_prob(torch.zeros(1)) == 0.0
    with pytest.raises(NotImplementedError):
        bnn.Parameter().sample()
    with pytest.raises(NotImplementedError):
        bnn.Model()(1)
    g = bnn.Gaussian(torch.Size((3, 2)))
    assert g.mu.dtype == torch.float32 and set(dict(g.named_parameters())) == {"mu", "rho", "zero", "one"}


def test_compat_log_prob_matches_reference_values():
    # the torch-op compatibility surface (not the hot path) on the KAT fixtures
    g = load_golden("gaussian.npz")
    q = bnn.Gaussian(torch.Size((3,)))
    q.mu.data, q.rho.data = torch.from_numpy(g["kat_mu"]), torch.from_numpy(g["kat_rho"])
    assert torch.equal(q.sigma, torch.from_numpy(g["kat_sigma"]))
    assert torch.equal(q.log_prob(torch.from_numpy(g["kat_w"])).detach(), torch.from_numpy(g["kat_logq"]))
    m = load_golden("mixture.npz")
    assert torch.equal(bnn.DEFAULT_SCALED_GAUSSIAN_MIXTURE.log_prob(torch.from_numpy(m["rnd_w"])).detach(),
                       torch.from_numpy(m["rnd_logp"]))
    assert bnn.DEFAULT_SCALED_GAUSSIAN_MIXTURE.sample() == 0.0


def test_all_layers_registry_converts_embedding_and_layernorm():
    net = torch.nn.Sequential(torch.nn.Embedding(11, 8, padding_idx=0), torch.nn.LayerNorm(8), torch.nn.Linear(8, 2))
    bm = bf.to_bayesian(net, delta=0.05, freeze=True, layers=bnn.TORCH2BAYE_ALL)
    kinds = [type(m).__name__ for m in bm.bayesian_children]
    assert kinds == ["Embedding", "LayerNorm", "Linear"]
    ln = bm.model[1]
    # gamma = 1 -> sigma = delta (to ~1e-5); beta = 0 -> rho = 0 (sigma = ln 2)
    assert torch.allclose(ln.weight.sigma, torch.full((8,), 0.05), rtol=1e-4)
    assert torch.all(ln.bias.rho == 0)
    assert bm.model[0].padding_idx == 0


def test_cpu_forward_fails_loudly():
    layer = bnn.Linear(4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.randn(2, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        bnn.Gaussian(torch.Size((4,))).sample()


def test_mc_samples_context_and_stream_ids():
    assert bf.runtime.get_mc_samples() == 1
    with bf.mc_samples(4):
        assert bf.runtime.get_mc_samples() == 4
        with bf.mc_samples(2):
            assert bf.runtime.get_mc_samples() == 2
        assert bf.runtime.get_mc_samples() == 4
    assert bf.runtime.get_mc_samples() == 1
    a, b = bnn.Gaussian(torch.Size((2,))), bnn.Gaussian(torch.Size((2,)))
    assert a.tensor_id != b.tensor_id
    bf.manual_seed(77)
    s0, s1 = a.next_stream(), a.next_stream()
    assert (s0.seed, s0.tensor_id) == (77, a.tensor_id) and s1.step == s0.step + 1

    class Fixed:
        def sample(self, size):
            return torch.ones(size)

    a.normal = Fixed()
    assert a.next_stream(3).eps.shape == (3, 2)



# ---------------------------------------------------------------- host logic of the extensions (no kernels run)
def test_presampler_tables_cover_every_element_once():
    """Chunk table / slot ranges of the multi-tensor sampling launch, built on the CPU: every quad of every
    tensor appears in exactly one chunk, chunks of one layer (weight + bias) are contiguous."""
    from bayeformers_b200.presample import Presampler
    net = torch.nn.Sequential(torch.nn.Linear(64, 136), torch.nn.Linear(136, 4104, bias=False), torch.nn.Linear(4104, 10))
    bm = bf.to_bayesian(net, delta=0.05, gemm_dtype="bf16")
    ps = Presampler(bm)
    S = 3
    tensors = ps._tensors(S)
    ps._build(S, tensors, torch.device("cpu"))
    cq = _lib.load().bf_sample_kl_multi_chunk_quads()
    chunks = ps.d_chunks.view(-1, 2).tolist()
    slots = ps.d_slots.view(-1, 2).tolist()
    assert len(chunks) == ps.n_chunks and len(slots) == ps.n_slots == 3
    covered = {}
    for ti, q0 in chunks:
        n = tensors[ti][1].mu.numel()
        nquad = max((n + 3) // 4, 1)
        assert q0 % cq == 0 and q0 < nquad
        covered.setdefault(ti, []).append(q0)
    for ti, (li, g, pr, dt) in enumerate(tensors):
        nquad = max((g.mu.numel() + 3) // 4, 1)
        assert covered[ti] == list(range(0, nquad, cq))
    assert slots[0][0] == 0 and slots[-1][1] == ps.n_chunks
    assert all(a[1] == b[0] for a, b in zip(slots, slots[1:]))
    # weights of tensor-core-eligible layers are drawn in bf16, the 10-wide head and all biases in fp32
    assert [t[3] for t in tensors] == [torch.bfloat16, torch.float32, torch.bfloat16, torch.float32, torch.float32]
    descs = (_lib.BfTensorDesc * len(tensors)).from_buffer_copy(bytes(ps.d_descs.numpy()))
    assert [d.n for d in descs] == [t[1].mu.numel() for t in tensors]
    offs = [d.w_out or 0 for d in descs]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs) and ps.arena_bytes >= offs[-1] + S * 10 * 4


def test_accelerate_host_swaps_layernorm_and_fuses_hf_gelu():
    from transformers import BertConfig, BertForSequenceClassification
    cfg = BertConfig(num_labels=2, num_hidden_layers=1, hidden_size=64, num_attention_heads=4, intermediate_size=128)
    model = BertForSequenceClassification(cfg)
    bm = bf.to_bayesian(model, delta=0.05, freeze=True)
    keys = list(bm.state_dict())
    bf.accelerate_host_(bm)
    layer = bm.model.bert.encoder.layer[0]
    assert type(layer.output.LayerNorm).__name__ == "HostLayerNorm" and isinstance(layer.output.LayerNorm, torch.nn.LayerNorm)
    assert layer.intermediate.dense.activation == "gelu"
    assert isinstance(layer.intermediate.intermediate_act_fn, torch.nn.Identity)
    assert layer.output.dense.activation is None  # only `dense` + activation blocks are touched
    assert list(bm.state_dict()) == keys  # checkpoint names unchanged
    # a ReLU block is left alone
    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.dense = torch.nn.Linear(8, 8)
            self.intermediate_act_fn = torch.nn.ReLU()
    other = bf.accelerate_host_(bf.to_bayesian(torch.nn.Sequential(Block())))
    assert other.model[0].dense.activation is None and isinstance(other.model[0].intermediate_act_fn, torch.nn.ReLU)


def test_accelerate_host_fuses_output_blocks():
    import copy
    from transformers import BertConfig, BertForSequenceClassification
    cfg = BertConfig(num_labels=2, num_hidden_layers=2, hidden_size=64, num_attention_heads=4, intermediate_size=128)
    bm = bf.to_bayesian(BertForSequenceClassification(cfg), delta=0.05, freeze=True)
    keys = list(bm.state_dict())
    bf.accelerate_host_(bm, fuse_residual=True)
    layer = bm.model.bert.encoder.layer[1]
    for blk in (layer.attention.output, layer.output):
        assert type(blk).__name__.startswith("Fused") and isinstance(blk, torch.nn.Module)
        assert type(blk).__mro__[2].__name__ in ("BertSelfOutput", "BertOutput")
    sites = [m._bf_site for m in bm.modules() if hasattr(m, "_bf_site")]
    assert len(sites) == 4 and len(set(sites)) == 4  # one dropout stream per block
    assert list(bm.state_dict()) == keys  # checkpoint names unchanged
    assert type(copy.deepcopy(bm).model.bert.encoder.layer[0].output).__name__ == "FusedBertOutput"
    # a block whose dense is not Bayesian, or with extra children, is left alone
    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.dense, self.LayerNorm, self.dropout = torch.nn.Linear(8, 8), torch.nn.LayerNorm(8), torch.nn.Dropout(0.1)
    plain = bf.accelerate_host_(torch.nn.Sequential(Block()), fuse_residual=True)
    assert type(plain[0]) is Block
    # CPU tensors: the stock forward of the block runs up to the Bayesian Linear, which refuses (no CPU path)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer.output(torch.randn(2, 128), torch.randn(2, 64))


def test_grad_sink_registry_and_guard():
    """runtime.GradSink bookkeeping (no kernels): one sink per input tensor and forward, reset by bnn.Model.forward,
    and the tensor hook that refuses topologies where in-place accumulation would lose a contribution."""
    from bayeformers_b200 import runtime
    runtime.reset_sinks()
    x = torch.randn(4, 8, requires_grad=True) * 2  # non-leaf, like a layer input
    assert runtime.sink_for(x, create=False) is None
    s1 = runtime.sink_for(x, create=True)
    assert runtime.sink_for(x, create=False) is s1 and runtime.sink_for(x.clone(), create=False) is None
    # valid case: the gradient autograd delivers for x IS the buffer the Linear layers accumulated into
    buf = torch.zeros(4, 8)
    s1.buffer, s1.used = buf, True
    s1.check(buf)  # no error; one-shot state is cleared
    assert s1.buffer is None and s1.used is False and s1.tensor is None
    # invalid case: some other consumer made the engine build a new sum after the accumulation
    s1.buffer, s1.used = buf, True
    with pytest.raises(RuntimeError, match="gradient sinks are not valid"):
        s1.check(buf + 1)
    # unused sinks never complain (inputs of layers that have no fused residual consumer)
    s2 = runtime.sink_for(torch.randn(2, 2, requires_grad=True) + 0, create=True)
    s2.check(torch.zeros(2, 2))
    runtime.reset_sinks()
    assert runtime.sink_for(x, create=False) is None
    # the hook is wired: backward through x calls check (no buffer set -> silent)
    runtime.enable_grad_sinks(False)
    y = torch.randn(3, 3, requires_grad=True) * 1.0
    s3 = runtime.sink_for(y, create=True)
    y.sum().backward()
    assert s3.tensor is None  # check() ran
    with pytest.raises(ValueError, match="fuse_residual"):
        bf.accelerate_host_(torch.nn.Sequential(torch.nn.Linear(2, 2)), fuse_residual=False, grad_sinks=True)
    assert runtime.grad_sinks_enabled() is False


def test_harness_fold_pick_and_predictive_stats():
    from bayeformers_b200 import harness
    x = torch.arange(6).view(2, 3)
    assert torch.equal(harness._fold(x, 3), torch.cat([x, x, x]))
    assert harness._fold("keep", 3) == "keep"
    out = {"a": torch.ones(2), "b": torch.zeros(2)}
    assert harness._pick(out, "a") is out["a"]
    assert harness._pick((1, 2, 3), (-2, -1)) == (2, 3)
    raw = torch.tensor([[[2.0, 0.0], [0.0, 1.0]], [[1.0, 0.0], [2.0, 0.0]]])  # [S=2, B=2, C=2]
    st = bf.predictive_stats(raw, torch.tensor([0, 1]))
    assert st["probs"].shape == (2, 2) and torch.allclose(st["probs"].sum(-1), torch.ones(2))
    assert float(st["acc"]) == 0.75 and abs(float(st["acc_std"]) - 0.25) < 1e-6


def test_extensions_refuse_cpu_tensors():
    with pytest.raises(RuntimeError, match="CUDA"):
        bf.optim.ClipAdamW([torch.nn.Parameter(torch.zeros(4))])
    ln = bf.accelerate_host_(torch.nn.Sequential(torch.nn.LayerNorm(256)))[0]
    y = ln(torch.randn(3, 256))  # CPU input: the stock implementation runs (host plumbing, not the variational path)
    assert y.shape == (3, 256)

# ---------------------------------------------------------------- C-ABI boundary
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bayeformers_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bf_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    from bayeformers_b200.build import build
    build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/bayeformers_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding and header disagree"
    assert _lib.load().bf_abi_version() == _lib.ABI_VERSION == 2


@pytest.mark.skipif(not os.path.isdir("/root/reference/bayeformers"), reason="live reference only in the build container")
@pytest.mark.parametrize("kw", [dict(delta=0.05, freeze=True), dict(delta=0.1, freeze=False), dict()])
def test_reference_checkpoint_loads_strictly_and_back(kw):
    """Drop-in checkpoint compatibility: a state_dict saved by the unmodified reference
    (examples/bert_glue.py:303-309 saves `b_model.state_dict()`) loads into this package's model with strict=True and
    vice versa -- same keys, same shapes, same values after the round trip."""
    import sys
    sys.path.insert(0, "/root/reference")
    try:
        from bayeformers import to_bayesian as ref_to_bayesian
    finally:
        sys.path.remove("/root/reference")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(10, 12), torch.nn.Tanh(), torch.nn.Linear(12, 3, bias=False))
    torch.manual_seed(5)
    ref = ref_to_bayesian(net, **kw)
    torch.manual_seed(6)  # different init draws on purpose: loading must overwrite them
    ours = bf.to_bayesian(net, **kw)
    ref_sd = ref.state_dict()
    assert list(ref_sd) == list(ours.state_dict())
    ours.load_state_dict(ref_sd, strict=True)
    for k, v in ours.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    back = ref_to_bayesian(net, **kw)
    back.load_state_dict(ours.state_dict(), strict=True)
    for k, v in back.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k


def test_rng_state_round_trip_restores_stream_ids_and_counters():
    """bf.rng_state / bf.load_rng_state (reproducible resume): seed, dropout salt, per-tensor stream ids / steps and the
    fused blocks' call counters survive a round trip into freshly built objects; pure host logic, no kernels."""
    import copy
    import bayeformers_b200 as bf
    import bayeformers_b200.nn as bnn
    from bayeformers_b200 import runtime

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 4))
    bf.manual_seed(4711)
    a = bf.to_bayesian(net, delta=0.05)
    gs = [m for m in a.modules() if isinstance(m, bnn.Gaussian)]
    for k, g in enumerate(gs):
        g.step = 10 + k
    runtime.set_dropout_salt(3)
    state = copy.deepcopy(bf.rng_state(a))
    assert state["seed"] == 4711 and state["dropout_salt"] == 3 and len(state["gaussians"]) == len(gs)
    bf.manual_seed(1)
    runtime.set_dropout_salt(0)
    b = bf.to_bayesian(net, delta=0.05)  # fresh stream ids
    gb = [m for m in b.modules() if isinstance(m, bnn.Gaussian)]
    assert [g.tensor_id for g in gb] != [g.tensor_id for g in gs]
    bf.load_rng_state(b, state)
    assert runtime.seed() == 4711 and runtime.dropout_seed() != runtime.seed()
    assert [(g.tensor_id, g.step) for g in gb] == [(g.tensor_id, g.step) for g in gs]
    runtime.set_dropout_salt(0)
    assert runtime.dropout_seed() == runtime.seed()
    with pytest.raises(KeyError):
        bf.load_rng_state(torch.nn.Sequential(torch.nn.Linear(2, 2)), state)


def test_native_attention_glue_registers_and_hands_bias_boxes():
    """accelerate_host_(attention=True) points a HuggingFace config at the native attention function and gives every
    self-attention block its dropout stream id; attention_bias_grads=True additionally installs the pre-hook that hands
    the Bayesian query / key / value projections a bias_grad_box each (host logic only: no kernel runs on CPU)."""
    transformers = pytest.importorskip("transformers")
    import bayeformers_b200 as bf
    from bayeformers_b200.nn.layers import attention as A

    cfg = transformers.BertConfig(vocab_size=50, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                  intermediate_size=256, max_position_embeddings=32, num_labels=2)
    for flag in (False, True):
        m = bf.to_bayesian(transformers.BertForSequenceClassification(cfg), delta=0.05, freeze=True)
        m = bf.accelerate_host_(m, layernorm=False, fuse_gelu=False, attention=True, attention_bias_grads=flag)
        blocks = [mod for mod in m.modules() if type(mod).__name__.endswith("SelfAttention")]
        assert len(blocks) == 2 and all(hasattr(b, "_bf_site") for b in blocks)
        assert len({b._bf_site for b in blocks}) == 2
        assert all((len(b._forward_pre_hooks) == 1) == flag for b in blocks)
        assert [c._attn_implementation for c in {id(x.config): x.config for x in m.modules() if hasattr(x, "config")}.values()] \
            == [A.NAME]
        if flag:
            A._qkv_pre_hook(blocks[0], ())
            boxes = blocks[0]._bf_qkv_boxes
            assert [blocks[0].query._bias_grad_box, blocks[0].key._bias_grad_box, blocks[0].value._bias_grad_box] == boxes
            assert all(b == [] for b in boxes) and boxes[0] is not boxes[1]


def test_gelu_epilogue_polynomials_meet_their_stated_error():
    """The fused GELU / GELU' epilogues (csrc/bf_gemm_act.cu: gelu_poly2, gelu_grad_poly2) evaluate odd polynomials on
    the argument clamped to [-4, 4].  Read the coefficients out of the CUDA source, evaluate them the way the kernel does
    (float32 Horner in z^2), and check the error bounds the source and DESIGN.md state against the erf forms in float64."""
    import math
    import os
    import re

    src = open(os.path.join(os.path.dirname(__file__), "..", "bayeformers_b200", "csrc", "bf_gemm_act.cu")).read()

    def coefficients(fn):
        body = src[src.index(f"bf_f2 {fn}(bf_f2 z2)"):]
        body = body[:body.index("\n}\n")]
        vals = [float(v) for v in re.findall(r"bf_splat2\((-?[0-9.]+e[-+][0-9]+)f\)", body)]
        return vals  # highest degree first, as the Horner chain lists them

    erf = np.vectorize(math.erf)
    z = np.linspace(-8.0, 8.0, 400001).astype(np.float32)
    zc = np.clip(z, np.float32(-4.0), np.float32(4.0))
    t = zc * zc
    z64 = z.astype(np.float64)
    phi = 0.5 * (1.0 + erf(z64 / math.sqrt(2.0)))
    for fn, n_coef, truth, tol_inside, tol_all in (
            ("gelu_poly2", 8, phi, 3e-5, 6e-5),
            ("gelu_grad_poly2", 9, phi + z64 * np.exp(-0.5 * z64 * z64) / math.sqrt(2.0 * math.pi), 1e-4, 6e-4)):
        c = coefficients(fn)
        assert len(c) == n_coef, (fn, c)
        acc = np.full_like(t, np.float32(c[0]))
        for ck in c[1:]:
            acc = acc * t + np.float32(ck)
        got = (np.float32(0.5) + zc * acc).astype(np.float64)  # Phi(z) resp. gelu'(z)
        err = np.abs(got - truth)
        assert err[np.abs(z) <= 4.0].max() < tol_inside, (fn, err[np.abs(z) <= 4.0].max())
        assert err.max() < tol_all, (fn, err.max())
    # and the activation itself: z * Phi(z) against the exact GELU, relative to bf16's rounding step at that magnitude
    c = coefficients("gelu_poly2")
    acc = np.full_like(t, np.float32(c[0]))
    for ck in c[1:]:
        acc = acc * t + np.float32(ck)
    y = (z * (np.float32(0.5) + zc * acc)).astype(np.float64)
    assert np.abs(y - z64 * phi).max() < 4.5e-4  # worst at |z| = 8 (clamped tail); bf16 rounding of 8.0 is 3.1e-2
