#!/usr/bin/env python
"""bench.py -- benchmarks of the variational-layer hot path, one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W                  # this repo (CUDA, sm_100a), headline config
    python bench.py --impl reference --gpus N --steps K ...        # the reference's own CPU path, same config
    python bench.py --config {bert_cls,bert_qa,bert_large,mlp,linear} [--shard {batch,samples}]

Workloads = BASELINE.json `configs` (synthetic inputs, random-init models):

  bert_cls   (default, the one the metric is quoted on) BERT-base sequence classification,
             to_bayesian(delta=0.05, freeze=True), T=128, S=4   -- pattern of examples/bert_glue.py:56-73,225-241
  bert_qa    BERT-base span head, T=384, S=8                     -- examples/bert_squad.py:190-212,216-234
  bert_large BERT-large, every Linear + Embedding + LayerNorm Bayesian (TORCH2BAYE_ALL), T=512, S=16
  mlp        MLP 784-512-10, batch 64, S=1                       -- examples/mlp_mnist.py:16-26
  linear     bnn.Linear 4096x4096, batch 8192, sweep S=1..32 (fwd+bwd, sample+KL GB/s)

A step = S-sample forward + ELBO loss + backward + grad clip + AdamW.  `--shard batch` (default): every rank holds all S
samples of its own sequences, identical Philox weights on every rank, NCCL all-reduce of the gradients.  `--shard
samples`: the S samples are split over the ranks (S % N == 0), every rank sees the whole global batch; logits are
averaged across ranks before the loss, gradients summed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")

import torch  # noqa: E402

N_BATCHES = 1000  # the reference divides the KL term by len(train_loader) (bert_glue.py:235)

# name -> (metric, unit, default T, default S, default per-GPU batch, reference batch)
CONFIGS = {
    "bert_cls": ("bayes_bert_base_train_seqs_per_s", "seq/s", 128, 4, 512, 8),
    "bert_qa": ("bayes_bert_base_qa_train_seqs_per_s", "seq/s", 384, 8, 64, 2),
    "bert_large": ("bayes_bert_large_all_bayesian_train_seqs_per_s", "seq/s", 512, 16, 8, 1),
    "mlp": ("bayes_mlp_784_512_10_train_imgs_per_s", "img/s", 0, 1, 64, 64),
    "linear": ("bayes_linear_4096_fwd_bwd_rows_per_s", "rows/s", 0, 4, 8192, 8192),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="bert_cls", choices=sorted(CONFIGS))
    ap.add_argument("--shard", default="batch", choices=["batch", "samples"])
    ap.add_argument("--batch", type=int, default=0, help="units per GPU per step (before the S-fold); 0 = config default "
                                                         "(bert_cls: 512 sequences -> 89 GB of HBM)")
    ap.add_argument("--graph", type=int, default=1, help="1: capture the whole training step in one CUDA graph; 0: eager")
    ap.add_argument("--samples", type=int, default=0, help="MC samples S; 0 = config default")
    ap.add_argument("--seq", type=int, default=0, help="sequence length; 0 = config default")
    ap.add_argument("--gemm", default="bf16", choices=["bf16", "fp32", "fp32x3"],
                    help="bf16: tcgen05 (1e-2 mode); fp32x3: reference precision on the tensor cores (3-pass bf16 split, "
                         "1e-5 mode); fp32: FFMA parity kernels")
    ap.add_argument("--kl-grad", type=int, default=1)
    ap.add_argument("--ref-batch", type=int, default=0, help="units per step of the CPU reference (0 = config default: "
                                                             "bert_cls 8 = BATCH_SIZE of examples/bert_glue.py:78)")
    ap.add_argument("--ref-samples", type=int, default=0, help="MC samples of the CPU reference (0 = the config's S, "
                                                               "except bert_large: 2, scaled linearly and said so)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", type=int, default=1, help="N=1 only: also time the fp32 parity mode and the reference's "
                                                          "own batch size (side-by-side rows of the JSON line)")
    ap.add_argument("--fused-optim", type=int, default=1, help="bf.optim.ClipAdamW instead of clip_grad_norm_ + AdamW")
    ap.add_argument("--presample", type=int, default=1, help="one multi-tensor sample+KL launch per forward")
    ap.add_argument("--fuse-gelu", type=int, default=1)
    ap.add_argument("--host-ln", type=int, default=1)
    ap.add_argument("--fuse-residual", type=int, default=1)
    ap.add_argument("--grad-sinks", type=int, default=1)
    ap.add_argument("--attention", type=int, default=1,
                    help="1 (default): the host model's short-sequence attention (T <= 128) runs on the native kernels "
                         "(tcgen05 at T = 128: 1.2 ms per BERT-base layer fwd+bwd against 2.2 ms of cuDNN's fused "
                         "attention); 0: torch SDPA")
    ap.add_argument("--attention-bias-grads", type=int, default=0,
                    help="1: q / k / v bias gradients from the attention backward kernel (no net gain measured: see "
                         "bf.accelerate_host_)")
    ap.add_argument("--resln-keep-bits", type=int, default=1,
                    help="1 (default): the fused dropout+residual+LayerNorm forward hands the mask's keep bits to its backward "
                         "(128 bytes per row); 0: the backward regenerates them from the Philox counter")
    ap.add_argument("--gelu-poly", type=int, default=1,
                    help="1 (default): the fused GELU / GELU' epilogues evaluate odd polynomials (|err| <= 1e-4 / 6e-4, "
                         "inside bf16 rounding); 0: the erf forms (bf_set_option(BF_OPT_GELU_POLY))")
    ap.add_argument("--sigma-cache", type=int, default=1,
                    help="ClipAdamW writes softplus(updated rho) next to its update; the sampling kernel reads it")
    ap.add_argument("--gelu-links", type=int, default=1,
                    help="fold GELU' into the dgrad epilogue of the Linear that consumes a fused-GELU layer's output")
    ap.add_argument("--layers", type=int, default=0, help="debug: override num_hidden_layers")
    ap.add_argument("--profile", action="store_true",
                    help="for runs under ncu only: allow < 3 warm-up steps, skip the e2e and CPU legs (numbers invalid)")
    a = ap.parse_args()
    metric, unit, T, S, B, RB = CONFIGS[a.config]
    a.metric, a.unit = metric, unit
    a.seq = a.seq or T
    a.samples = a.samples or S
    a.batch = a.batch or B
    a.ref_batch = a.ref_batch or RB
    if not a.ref_samples:
        a.ref_samples = 2 if a.config == "bert_large" else a.samples
    return a


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            p.update(json.load(open(f)))
            p["source"] = "measured"
        except Exception:
            pass
    return p


# --------------------------------------------------------------------------- workloads
class Workload:
    """Model + synthetic inputs + loss of one BASELINE.json config; shared by the GPU arm and the CPU reference arm."""

    def __init__(self, args):
        self.args, self.name = args, args.config

    # ---- frequentist model (random init; HF zero-inits biases: perturb so MOPED sees non-degenerate values)
    def build(self):
        a = self.args
        torch.manual_seed(0)
        if self.name == "mlp":
            model = torch.nn.Sequential(torch.nn.Linear(784, 512), torch.nn.ReLU(), torch.nn.Linear(512, 10))
            return model, None
        from transformers import BertConfig, BertForQuestionAnswering, BertForSequenceClassification
        if self.name == "bert_large":
            cfg = BertConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                             num_labels=2)
        else:
            cfg = BertConfig(num_labels=2)
        if a.layers:
            cfg.num_hidden_layers = a.layers
        model = BertForQuestionAnswering(cfg) if self.name == "bert_qa" else BertForSequenceClassification(cfg)
        with torch.no_grad():
            g = torch.Generator().manual_seed(1)
            for n, p in model.named_parameters():
                if n.endswith("bias"):
                    p.add_(torch.randn(p.shape, generator=g) * 0.02)
        return model, cfg

    def inputs(self, cfg, B, rank=0):
        """Host tensors of one step: dict(model inputs), tuple(targets)."""
        a = self.args
        g = torch.Generator().manual_seed(100 + rank)
        if self.name == "mlp":
            return {"x": torch.rand(B, 784, generator=g)}, (torch.randint(0, 10, (B,), generator=g),)
        ids = torch.randint(0, cfg.vocab_size, (B, a.seq), generator=g)
        if self.name == "bert_qa":
            return {"input_ids": ids}, (torch.randint(0, a.seq, (B,), generator=g), torch.randint(0, a.seq, (B,), generator=g))
        return {"input_ids": ids}, (torch.randint(0, 2, (B,), generator=g),)

    def forward(self, model, inp):
        """-> tuple of output tensors whose leading dimension is the (folded) batch."""
        if self.name == "mlp":
            return (model(inp["x"]),)
        out = model(**inp)
        if self.name == "bert_qa":
            return (out.start_logits, out.end_logits)
        return (out.logits,)

    @staticmethod
    def nll(means, targets):
        """Loss of the sample-MEAN predictions (bert_glue.py:69,234; bert_squad.py: mean start / end logits)."""
        ce = torch.nn.functional.cross_entropy
        return sum(ce(m.float(), t) for m, t in zip(means, targets)) / len(means)

    def flops_per_unit_sample(self, cfg):
        """fwd+bwd matmul flops of one unit (sequence / image) for ONE MC sample: (linear, attention)."""
        if self.name == "mlp":
            return 6.0 * (784 * 512 + 512 * 10), 0.0
        T = self.args.seq
        H, L, FF = cfg.hidden_size, cfg.num_hidden_layers, cfg.intermediate_size
        head = H * (2 if self.name == "bert_qa" else cfg.num_labels) + (0 if self.name == "bert_qa" else H * H)
        lin = L * (4 * H * H + 2 * H * FF) + head
        return 3.0 * (2 * T * lin), 3.0 * (L * 4 * T * T * H)

    def describe(self):
        a = self.args
        return {"bert_cls": "BERT-base to_bayesian(delta=0.05, freeze=True) GLUE-style classification, synthetic tokens",
                "bert_qa": "BERT-base to_bayesian(delta=0.05, freeze=True) SQuAD-style span head, synthetic tokens",
                "bert_large": "BERT-large to_bayesian(delta=0.05, freeze=True, all Linear + Embedding + LayerNorm), "
                              "synthetic tokens",
                "mlp": "Bayesian MLP 784-512-10 (to_bayesian(delta=0.05)), synthetic 28x28 inputs",
                "linear": "bnn.Linear 4096x4096 (default init, scale-mixture prior), x ~ N(0,1)"}[self.name] + \
            ", training step (S-sample fwd + ELBO + bwd + clip + AdamW)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 7:
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- reference arm / cpu baseline
def load_reference():
    """The UNMODIFIED reference package from oracle/_ref (a copy of /root/reference/bayeformers made by
    oracle/build_ref.py in the build container; git-ignored, travels to the GPU box).  None when absent."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "bayeformers")):
        return None
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import bayeformers  # noqa: F401
        import bayeformers.nn  # noqa: F401
        return bayeformers
    except Exception as e:  # pragma: no cover
        sys.stderr.write(f"[bench] oracle/_ref present but not importable ({type(e).__name__}: {e}); using the port\n")
        return None


def cpu_reference_steps(args, steps: int, warmup: int):
    """The reference's CPU path for this workload on all host cores, on a bounded sample of `--ref-batch` units and
    `--ref-samples` MC samples per step: the reference package itself (oracle/_ref: to_bayesian + the S-loop of
    examples/bert_glue.py:56-73 with torch.optim.AdamW -- transformers' AdamW, bert_glue.py:13, no longer exists)
    when it is present, else the oracle port of the same operator sequence (oracle/bayes_oracle.py).  bert_large
    (Embedding / LayerNorm Bayesian too) has no reference module for those layers: always the composed oracle.
    Returns dict(value, s_per_step, cores, kind, batch, samples)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = Workload(args)
    B, S = args.ref_batch, args.ref_samples
    if args.config == "linear":
        return cpu_reference_linear(args, steps, warmup)
    model, cfg = wl.build()
    ref = load_reference() if args.config != "bert_large" else None
    torch.manual_seed(2)
    if ref is not None:
        bm = ref.to_bayesian(model, delta=0.05, freeze=(args.config != "mlp")).train()
        log_prior, log_q = bm.log_prior, bm.log_variational_posterior
        kind = "reference"
    else:
        from oracle import bayes_oracle as O
        bm = O.oracle_convert(model, delta=0.05, freeze=(args.config != "mlp"), all_layers=(args.config == "bert_large")).train()
        log_prior, log_q = (lambda: O.model_log_prior(bm)), (lambda: O.model_log_variational_posterior(bm))
        kind = "port"
    inp, targets = wl.inputs(cfg, B)
    params = [p for p in bm.parameters() if p.requires_grad]
    optim = torch.optim.AdamW(params, lr=2e-5, eps=1e-8)

    def step():  # same work as the GPU step: S-loop fwd, ELBO, bwd, clip, AdamW (bert_glue.py:230-241)
        optim.zero_grad(set_to_none=True)
        outs, lps, lqs = [], [], []
        for _ in range(S):
            outs.append(wl.forward(bm, inp))
            lps.append(torch.as_tensor(log_prior()))
            lqs.append(torch.as_tensor(log_q()))
        means = [torch.stack([o[i] for o in outs]).mean(0) for i in range(len(outs[0]))]
        loss = (torch.stack(lqs).mean() - torch.stack(lps).mean()) / N_BATCHES + wl.nll(means, targets)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        optim.step()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    # per-unit throughput at the config's S: the S-loop is linear in S (one forward per sample)
    value = B / (dt * args.samples / S)
    return {"value": value, "s_per_step": dt, "cores": cores, "kind": kind, "batch": B, "samples": S}


def cpu_reference_linear(args, steps, warmup):
    cores = os.cpu_count() or 1
    ref = load_reference()
    S = min(args.ref_samples, 2)
    torch.manual_seed(0)
    x = torch.randn(args.ref_batch, 4096, requires_grad=True)
    if ref is not None:
        layer, kind = ref.nn.Linear(4096, 4096), "reference"
    else:
        from oracle import bayes_oracle as O
        layer, kind = O.oracle_convert(torch.nn.Sequential(torch.nn.Linear(4096, 4096)), None, False)[0], "port"

    def step():
        for _ in range(S):
            layer(x).square().mean().backward()
        layer.zero_grad(set_to_none=True)
        x.grad = None

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"value": args.ref_batch / (dt * args.samples / S), "s_per_step": dt, "cores": cores, "kind": kind,
            "batch": args.ref_batch, "samples": S}


def reference_config(args, r):
    """What the reference arm ACTUALLY ran (its own batch, dtype and optimizer -- not the GPU arm's)."""
    wl = Workload(args)
    scaled = "" if r["samples"] == args.samples else f" (timed with S={r['samples']}, scaled linearly to S={args.samples}: the S-loop is one forward per sample)"
    return {"workload": wl.describe(), "seq_len": args.seq, "mc_samples": args.samples, "batch_per_gpu": r["batch"],
            "global_batch": r["batch"], "gemm": "fp32 (torch CPU, MKL)", "kl_grad": False,
            "optimizer": "clip_grad_norm_ + torch.optim.AdamW", "sampling": "per layer, per sample (the reference's S-loop)" + scaled,
            "implementation": "unmodified reference package (oracle/_ref)" if r["kind"] == "reference"
                              else "oracle port of the reference's operator sequence (oracle/bayes_oracle.py)",
            "parallelism": "1 process, all host cores", "l2": "n/a (CPU)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_steps(args, args.steps, args.warmup)
    sample = (f"{r['batch']} units x S={r['samples']} per step (bounded sample of the workload), {r['kind']}, torch CPU fp32, "
              f"{r['cores']} threads, {r['s_per_step']:.2f} s/step")
    line = {"impl": "reference", "metric": args.metric, "value": r["value"], "unit": args.unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": reference_config(args, r),
            "cpu_baseline": {"value": r["value"], "unit": args.unit, "cores": r["cores"], "kind": r["kind"], "sample": sample},
            "e2e": {"value": r["value"], "unit": args.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    wl = Workload(args)
    bert = args.config.startswith("bert")
    shard = ("batch sharded, identical Philox weights per rank, NCCL grad all-reduce" if args.shard == "batch" else
             f"MC samples sharded ({args.samples // world} per rank), whole global batch on every rank, logits averaged "
             "across ranks before the loss, gradients summed")
    global_batch = args.batch * world
    return {"workload": wl.describe(), "seq_len": args.seq, "mc_samples": args.samples, "batch_per_gpu": args.batch,
            "global_batch": global_batch, "gemm": args.gemm, "kl_grad": bool(args.kl_grad),
            "optimizer": "bf.optim.ClipAdamW (fused clip + AdamW)" if args.fused_optim else "clip_grad_norm_ + torch AdamW(fused)",
            "sampling": "multi-tensor (1 launch per forward)" if args.presample else "per layer",
            "ffn_gelu": (("fused into bnn.Linear (forward epilogue; GELU' in the consumer's dgrad epilogue)" if args.gelu_links
                          else "fused into bnn.Linear (forward epilogue + GELU'/bias-grad pass)")
                         if args.fuse_gelu else "torch") if bert else "n/a",
            "host_layernorm": ("native kernels (bf_layernorm_*)" if args.host_ln else "torch") if bert else "n/a",
            "output_blocks": ("dropout + residual + LayerNorm fused (bf_resln_*, Philox mask, bias grad handed to the Linear)"
                              if args.fuse_residual else "torch dropout + add, separate LayerNorm") if bert else "n/a",
            "attention": ("native whole-sequence kernels (bf_attention_*: tcgen05 at T = 128, Philox dropout mask) for T <= 128, else torch SDPA"
                          if args.attention else "torch SDPA (cuDNN)") if bert else "n/a",
            "shared_input_grads": ("accumulated in place by the dgrad kernels (TMA reduce-add)"
                                   if (args.grad_sinks and args.fuse_residual) else "autograd add passes") if bert else "n/a",
            "parallelism": f"dp{world} ({shard})",
            "l2": "working set (sampled weights + activations, GBs) far exceeds the 126 MB L2; no explicit flush"
                  if bert else "flushed: a 256 MB buffer is rewritten between timed steps"}


# --------------------------------------------------------------------------- our arm
def measure(args, dev, world, rank, local, *, batch, gemm, steps, warmup, timing=True, e2e=True):
    """Build the workload at `batch` units per GPU in `gemm` mode and time `steps` steps.  Returns a dict."""
    import torch.distributed as dist

    import bayeformers_b200 as bf
    import bayeformers_b200.nn as bnn
    from bayeformers_b200 import ops, parallel

    wl = Workload(args)
    bert = args.config.startswith("bert")
    model, cfg = wl.build()
    bf.manual_seed(1234)
    bf.runtime.enable_gelu_links(bool(args.gelu_links))
    from bayeformers_b200 import _lib as _bf_lib
    _bf_lib.load().bf_set_option(_bf_lib.BF_OPT_GELU_POLY, int(args.gelu_poly))
    ops.resln_keep_bits["on"] = bool(args.resln_keep_bits)
    layers = bnn.TORCH2BAYE_ALL if args.config == "bert_large" else None
    bm = bf.to_bayesian(model, delta=0.05, freeze=(args.config != "mlp"), gemm_dtype=gemm, kl_grad=bool(args.kl_grad),
                        layers=layers)
    if bert and (args.host_ln or args.fuse_gelu or args.fuse_residual or args.attention):
        # same parameters and numerics: native LayerNorm kernels (fp32 gamma/beta); FFN GELU fused into the layer;
        # dropout + residual + LayerNorm of the output blocks in one pass each way (Philox dropout mask)
        bf.accelerate_host_(bm, layernorm=bool(args.host_ln), fuse_gelu=bool(args.fuse_gelu),
                            fuse_residual=bool(args.fuse_residual),
                            grad_sinks=bool(args.grad_sinks and args.fuse_residual), attention=bool(args.attention),
                            attention_bias_grads=bool(args.attention_bias_grads))
    bm = bm.to(dev).train()
    if args.presample:
        bf.enable_presample(bm)
    S_total = args.samples
    sample_shard = args.shard == "samples" and world > 1
    S = S_total
    if world > 1:
        parallel.broadcast_seed(0)
        if sample_shard:
            S = parallel.shard_samples(S_total)  # re-keys this rank's eps stream
    if gemm == "bf16":
        # activations flow in bf16 (embeddings / LayerNorm of the host model cast to bf16; fp32 masters in ClipAdamW);
        # the variational masters (mu, rho, priors) stay fp32
        bf.cast_frequentist_(bm, torch.bfloat16)
    params = [p for p in bm.parameters() if p.requires_grad]
    use_graph = bool(args.graph) and not args.profile
    if args.fused_optim:  # global-norm clip + AdamW in two launches (section 8f row 3)
        # model=bm: the optimizer keeps a sigma = softplus(rho) cache current for the sampling kernel
        optim = bf.optim.ClipAdamW(params, lr=2e-5, eps=1e-8, weight_decay=0.01, max_grad_norm=1.0,
                                   model=bm if args.sigma_cache else None)
    else:
        optim = torch.optim.AdamW(params, lr=2e-5, eps=1e-8, fused=True, capturable=use_graph)
    sync = parallel.GradSync(bm, average=not sample_shard)
    bf.enable_device_step(dev)  # eps = f(seed, tensor, host_step + device_step, sample): graph replays draw fresh eps

    # batch sharding: own sequences per rank; sample sharding: every rank holds the same global batch
    B = batch * world if sample_shard else batch
    inp_host, tgt_host = wl.inputs(cfg, B, rank=0 if sample_shard else rank)
    act_dtype = torch.bfloat16 if gemm == "bf16" else torch.float32
    inp_host = {k: (v.to(act_dtype) if v.is_floating_point() else v).pin_memory() for k, v in inp_host.items()}
    tgt_host = tuple(t.pin_memory() for t in tgt_host)
    inp_dev = {k: v.to(dev) for k, v in inp_host.items()}
    tgt_dev = tuple(t.to(dev) for t in tgt_host)
    out_host = torch.zeros(3, dtype=torch.float32).pin_memory()
    h2d_bytes = sum(v.numel() * v.element_size() for v in inp_host.values()) + sum(t.numel() * t.element_size() for t in tgt_host)

    def step_body(inp, tgt):
        bf.advance_step()
        if sync.bucketed:
            sync.zero_grad()  # gradients live in flat all-reduce buckets: one memset each
        else:
            optim.zero_grad(set_to_none=True)
        with bf.mc_samples(S):
            outs = wl.forward(bm, {k: v.repeat(S, *([1] * (v.dim() - 1))) for k, v in inp.items()})
        raws = [o.float().view(S, B, *o.shape[1:]) for o in outs]
        lp, lq = bm.log_prior(), bm.log_variational_posterior()
        if sample_shard:
            means = [parallel.mean_over_samples(r, S_total) for r in raws]
            kl = (lq.sum() - lp.sum()) / S_total  # this rank's share; summed over ranks by the gradient all-reduce
        else:
            means = [r.mean(0) for r in raws]
            kl = lq.mean() - lp.mean()
        loss = kl / N_BATCHES + wl.nll(means, tgt)
        loss.backward()
        sync.finish()
        if not args.fused_optim:
            torch.nn.utils.clip_grad_norm_(params, 1.0)
        optim.step()
        return loss, lp.mean(), lq.mean()

    graph_note = "eager"
    step = step_body
    graph = None
    if use_graph:
        static_inp = {k: v.clone() for k, v in inp_dev.items()}
        static_tgt = tuple(t.clone() for t in tgt_dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            # every Linear contraction is ours, so nothing has touched cuBLAS yet; SDPA may fall back to
            # bmm under capture, and creating a cuBLAS handle while capturing is illegal: make it exist now
            _d = torch.ones(64, 64, device=dev, dtype=torch.bfloat16)
            torch.bmm(_d[None], _d[None]); torch.mm(_d.float(), _d.float())
            for _ in range(3):
                step_body(static_inp, static_tgt)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        optim.zero_grad(set_to_none=True) if not sync.bucketed else None
        torch.cuda.empty_cache()  # the eager warm-up's activation blocks: the graph gets its own pool
        graph = torch.cuda.CUDAGraph()
        if not sync.bucketed:
            optim.zero_grad(set_to_none=True)
        l0 = ops.stats["launches"]
        # bf.hf_capture_compat: keep HF on the mask-free fused-attention path while capturing (see its docstring)
        with bf.hf_capture_compat(), torch.cuda.graph(graph, stream=side):  # same stream as the warm-up
            static_out = step_body(static_inp, static_tgt)
        launches_per_graph = ops.stats["launches"] - l0

        def step(inp, tgt):  # noqa: F811
            if inp is not static_inp:
                for k in static_inp:
                    static_inp[k].copy_(inp[k], non_blocking=True)
                for d_, s_ in zip(static_tgt, tgt):
                    d_.copy_(s_, non_blocking=True)
            graph.replay()
            ops.stats["launches"] += launches_per_graph
            return static_out

        inp_dev, tgt_dev = static_inp, static_tgt
        graph_note = "whole training step captured in one CUDA graph"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = None if bert else torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def run_steps(n, e2e_mode):
        """n steps bracketed by events; small workloads get an L2 flush between steps (outside the event brackets)."""
        total = 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if flush is None:
            e0.record()
        for _ in range(n):
            if flush is not None:
                flush.fill_(1)
                e0.record()
            if e2e_mode:
                i_ = {k: v.to(dev, non_blocking=True) for k, v in inp_host.items()}
                t_ = tuple(t.to(dev, non_blocking=True) for t in tgt_host)
                loss, lp, lq = step(i_, t_)
                out_host.copy_(torch.stack([loss.detach().float(), lp.detach().float(), lq.detach().float()]), non_blocking=True)
                torch.cuda.synchronize()
            else:
                step(inp_dev, tgt_dev)
            if flush is not None:
                e1.record()
                torch.cuda.synchronize()
                total += e0.elapsed_time(e1)
        if flush is None:
            e1.record()
            torch.cuda.synchronize()
            total = e0.elapsed_time(e1)
        return total / max(n, 1)

    n_warm = warmup if args.profile else max(warmup, 3)
    for _ in range(n_warm):
        step(inp_dev, tgt_dev)

    # ---- timed region 1: inputs resident in HBM (value)
    ops.enable_kernel_timing(timing and not use_graph)
    launches0 = ops.stats["launches"]
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    if args.profile:  # `ncu --profile-from-start off`: only the measured step(s) are profiled
        ops.enable_nvtx(True)
        torch.cuda.profiler.start()
    ms = run_steps(steps, False)
    barrier()
    if args.profile:
        torch.cuda.profiler.stop()
    clk = clocks.stop() if clocks else None
    launches = ops.stats["launches"] - launches0
    kern = ops.kernel_timing_summary()
    ops.enable_kernel_timing(False)
    # ---- timed region 2: end to end through the public API with HOST buffers (e2e)
    ms_e2e = None
    if e2e and not args.profile:
        barrier()
        ms_e2e = run_steps(steps, True)
        barrier()

    kern_steps = steps
    if use_graph and timing:
        # CUDA events cannot bracket kernels inside a replayed graph: take the per-kernel durations from eager,
        # instrumented executions of the same step after the timed regions.  The graph (and its private memory pool,
        # half of the GPU at the default batch) is released first.
        static_out = None
        graph = None
        step = step_body
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        kern_steps = 2
        step_body(inp_dev, tgt_dev)
        barrier()
        ops.enable_kernel_timing(True)
        for _ in range(kern_steps):
            step_body(inp_dev, tgt_dev)
        barrier()
        kern = ops.kernel_timing_summary()
        ops.enable_kernel_timing(False)
    if world > 1:
        t = torch.tensor([ms, ms_e2e if ms_e2e is not None else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), (float(t[1]) if ms_e2e is not None else None)
    ar_ms = sync.allreduce_alone_ms() if world > 1 else None
    res = {"ms": ms, "ms_e2e": ms_e2e, "units_per_step": B if sample_shard else B * world, "kern": kern,
           "allreduce_alone_ms": ar_ms, "allreduce_launches": len(sync.buckets) if sync.bucketed else None,
           "kern_steps": kern_steps, "launches": launches, "clocks": clk, "graph_note": graph_note, "n_warm": n_warm,
           "h2d_bytes": h2d_bytes, "cfg": cfg, "hbm_peak_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
           "allreduce_bytes": sync.bytes_last_step, "use_graph": use_graph, "wl": wl}
    # tear down so that another measurement can follow in this process
    sync.remove()
    bf.disable_device_step()
    bf.runtime.enable_grad_sinks(False)
    del graph, bm, optim, model
    torch.cuda.empty_cache()
    return res


def sample_kl_sweep(args, dev, pk):
    """The multi-tensor sample+KL launch of THIS workload's model alone (no other kernel running), S = 1, 4, 16: CUDA
    events around each launch, algorithmic bytes as in SURVEY.md 8d.  The kernel is instruction-bound (Philox4x32-10 +
    Box-Muller per element-sample), so GB/s falls with S while element-samples/s rises; the in-step figure
    (`roofline_sample_kl`) is the same kernel at the power-capped clock of the step."""
    import bayeformers_b200 as bf
    import bayeformers_b200.nn as bnn
    from bayeformers_b200 import ops

    wl = Workload(args)
    model, _ = wl.build()
    bm = bf.to_bayesian(model, delta=0.05, freeze=(args.config != "mlp"), gemm_dtype=args.gemm,
                        layers=bnn.TORCH2BAYE_ALL if args.config == "bert_large" else None).to(dev)
    bf.enable_presample(bm)
    rows = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for S in (1, 4, 16):
        for _ in range(3):
            bm._presampler.run(S)
        torch.cuda.synchronize()
        ops.enable_kernel_timing(True)
        for _ in range(5):
            flush.fill_(1)
            bm._presampler.run(S)
        torch.cuda.synchronize()
        k = ops.kernel_timing_summary()["sample_kl_fwd"]
        ops.enable_kernel_timing(False)
        gbs = k["work"] / (k["ms"] / 1e3) / 1e9
        n_elem = sum(n for _, n, _ in bm._presampler.offsets)
        rows.append({"S": S, "us_per_launch": k["ms"] / k["calls"] * 1e3, "GB/s": gbs, "frac_of_hbm_peak": gbs / pk["hbm_gbs"],
                     "G_element_samples_per_s": n_elem * S / (k["ms"] / k["calls"] / 1e3) / 1e9,
                     "algorithmic_bytes_per_launch": k["work"] / k["calls"]})
    del bm, model
    torch.cuda.empty_cache()
    return rows


def traffic_record(kernel_key):
    """DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch) of the dominant kernel from a committed
    ncu capture AT THE BENCH SHAPE: profiles/traffic.json, written by scripts/ncu_traffic.py from an `ncu --set full`
    report.  None when no capture has been committed (never a hand-copied constant)."""
    f = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        rec = json.load(open(f)).get(kernel_key)
    except Exception:
        return None
    return rec


def run_linear_sweep(args, dev):
    """BASELINE.json configs[1]: bnn.Linear 4096x4096, batch 8192, S sweep, fwd+bwd; sample+KL GB/s per S."""
    import bayeformers_b200 as bf
    import bayeformers_b200.nn as bnn
    from bayeformers_b200 import ops

    pk = peaks()
    torch.manual_seed(0)
    bf.manual_seed(1)
    N = K = 4096
    B = args.batch
    rows = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for S in (1, 2, 4, 8, 16, 32):
        layer = bnn.Linear(K, N).to(dev)
        layer.gemm_dtype = bf.runtime._as_dtype(args.gemm)
        layer.kl_grad = True
        x = torch.randn(S * B, K, device=dev, dtype=torch.bfloat16 if args.gemm == "bf16" else torch.float32).requires_grad_()

        def step():
            with bf.mc_samples(S):
                y = layer(x)
            loss = y.float().square().mean() + 1e-6 * (layer.live_log_variational_posterior - layer.live_log_prior).mean()
            loss.backward()
            layer.zero_grad(set_to_none=True)
            x.grad = None

        for _ in range(max(args.warmup, 3)):
            step()
        torch.cuda.synchronize()
        it = max(2, min(args.steps, 5 if S <= 8 else 2))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0.0
        for _ in range(it):
            flush.fill_(1)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        ms = total / it
        ops.enable_kernel_timing(True)
        step()
        torch.cuda.synchronize()
        k = ops.kernel_timing_summary()
        ops.enable_kernel_timing(False)
        sk = k.get("sample_kl_fwd", {"ms": 0.0, "work": 0.0})
        g_ms = sum(v["ms"] for n, v in k.items() if n.startswith("gemm_"))
        g_fl = sum(v["work"] for n, v in k.items() if n.startswith("gemm_"))
        rows.append({"S": S, "ms_fwd_bwd": ms, "rows_per_s": B / (ms / 1e3), "layer_tflops": 6.0 * S * B * N * K / ms / 1e9,
                     "contractions_tflops": g_fl / max(g_ms, 1e-9) / 1e9,
                     "sample_kl_gbs": sk["work"] / max(sk["ms"], 1e-9) / 1e6,
                     "sample_kl_frac_of_hbm": sk["work"] / max(sk["ms"], 1e-9) / 1e6 / pk["hbm_gbs"]})
        del layer, x
        torch.cuda.empty_cache()
    head = next(r for r in rows if r["S"] == args.samples)
    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_reference_steps(args, 1, 1)
        cpu = {"value": r["value"], "unit": args.unit, "cores": r["cores"], "kind": r["kind"],
               "sample": f"{r['batch']} rows x S={r['samples']} (scaled linearly to S={args.samples}), 1 warm-up + 1 timed "
                         f"fwd+bwd ({r['s_per_step']:.2f} s)"}
    line = {"metric": args.metric, "value": head["rows_per_s"], "unit": args.unit, "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": head["ms_fwd_bwd"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.gemm == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": Workload(args).describe().replace(", training step (S-sample fwd + ELBO + bwd + clip + AdamW)",
                                                                      ", fwd + bwd (no optimizer), kl_grad on"),
                       "batch_per_gpu": B, "mc_samples": args.samples, "gemm": args.gemm,
                       "l2": "flushed: a 256 MB buffer is rewritten between timed iterations"},
            "roofline": {"bound": "tensor", "kernel": "tcgen05 contractions (fwd, dgrad, fused wgrad)",
                         "achieved": head["contractions_tflops"], "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": head["contractions_tflops"] / pk["bf16_tflops"],
                         "peak_source": f"{pk['source']} bf16_tflops (burst: kernels timed alone)", "traffic": None},
            "sweep": rows, "cpu_baseline": cpu, "e2e": None, "gpu_launches": None}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.shard == "samples" and args.samples % world != 0:
        raise SystemExit(f"--shard samples needs S ({args.samples}) to be a multiple of the number of GPUs ({world})")
    pk = peaks()
    if os.environ.get("BF_ANOMALY"):  # debugging aid: forward traceback of a failing backward node
        torch.autograd.set_detect_anomaly(True)
    if args.config == "linear":
        if rank == 0:
            run_linear_sweep(args, dev)
        return

    try:
        res = measure(args, dev, world, rank, local, batch=args.batch, gemm=args.gemm, steps=args.steps, warmup=args.warmup)
    except Exception as e:  # e.g. a host sync inside the host model under capture: restart this process in eager mode
        if not (args.graph and world == 1 and not args.profile):
            raise
        import traceback
        traceback.print_exc(file=sys.stderr)
        sys.stderr.write(f"[bench] run with CUDA-graph capture failed ({type(e).__name__}); re-running eagerly\n")
        sys.stderr.flush()
        os.execv(sys.executable, [sys.executable] + sys.argv + ["--graph", "0"])

    def leave():
        """End of a multi-rank run.  Tearing NCCL down while a CUDA graph that captured its collectives is still
        alive can hang at exit: leave without running destructors."""
        if world == 1:
            return
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)

    if rank != 0:
        leave()
        return

    ms, ms_e2e, kern, kern_steps, cfg, wl = res["ms"], res["ms_e2e"], res["kern"], res["kern_steps"], res["cfg"], res["wl"]
    units = res["units_per_step"]
    value = units / (ms / 1e3)
    e2e = units / (ms_e2e / 1e3) if ms_e2e else None
    # ---- roofline of the dominant kernel family (tcgen05 contractions), from the CUDA-event brackets
    gemm = {k: v for k, v in kern.items() if k.startswith("gemm_")}
    g_ms = sum(v["ms"] for v in gemm.values())
    g_flops = sum(v["work"] for v in gemm.values())
    g_calls = sum(v["calls"] for v in gemm.values())
    peak_tf = pk["bf16_tflops_sustained"]
    ach_tf = g_flops / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
    tr = traffic_record(f"{args.config}:gemm_fwd_ffn_up")
    roofline = {"bound": "tensor",
                "kernel": "tcgen05 contractions: tc2::bayes_gemm2_kernel (fwd, dgrad; cta_group::2) + wg::bayes_wgrad_kernel "
                          "(fused wgrad), all layers",
                "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                "peak_source": f"{pk['source']} bf16_tflops_sustained (kernel timed inside a long step; the peak is cuBLAS "
                               "running back to back for 4 s under the power cap, these launches are interleaved with "
                               "lighter kernels and can clock higher, so frac may slightly exceed 1)",
                "traffic": None if tr is None else tr.get("dram_bytes_per_launch"),
                "traffic_source": None if tr is None else tr,
                "launches_per_step": g_calls / kern_steps, "avg_launch_ms": g_ms / max(g_calls, 1),
                "share_of_step": g_ms / kern_steps / ms,
                "timing": "CUDA events around every launch on the launching stream" +
                          (", taken in eager executions of the same step after the timed graph replays" if res["use_graph"]
                           else ", inside the timed region")}
    sk = kern.get("sample_kl_fwd")
    roof_sk = None
    if sk and sk["ms"] > 0:
        gbs = sk["work"] / (sk["ms"] / 1e3) / 1e9
        roof_sk = {"bound": "hbm", "kernel": "sample_kl_multi_kernel" if args.presample else "sample_kl_fwd_fast_kernel",
                   "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                   "share_of_step": sk["ms"] / kern_steps / ms, "mc_samples": args.samples,
                   "algorithmic_bytes_per_launch": sk["work"] / max(sk["calls"], 1), "traffic": None}
    lin_f, att_f = wl.flops_per_unit_sample(cfg)
    step_tf = (lin_f + att_f) * args.samples * value / 1e12
    kernels = {k: {"calls_per_step": v["calls"] / kern_steps, "ms_per_step": v["ms"] / kern_steps,
                   "share": v["ms"] / kern_steps / ms} for k, v in sorted(kern.items())}

    cpu = None
    if world == 1 and not args.no_cpu_baseline and not args.profile:
        r = cpu_reference_steps(args, 2, 1)
        scaled = "" if r["samples"] == args.samples else f", scaled linearly to S={args.samples}"
        cpu = {"value": r["value"], "unit": args.unit, "cores": r["cores"], "kind": r["kind"],
               "sample": f"{r['batch']} units x S={r['samples']}{scaled}, 1 warm-up + 2 timed steps of the reference's S-loop "
                         f"({r['s_per_step']:.2f} s/step), torch CPU fp32"}

    extras = None
    if world == 1 and args.extras and not args.profile and args.config == "bert_cls" and args.gemm == "bf16":
        # side-by-side rows: (1) the reference-precision (fp32, 1e-5) mode, (2) the reference's own batch size
        extras = {}
        try:
            r32 = measure(args, dev, 1, 0, local, batch=128, gemm="fp32x3", steps=5, warmup=3, timing=False, e2e=False)
            extras["value_fp32"] = {"value": 128 / (r32["ms"] / 1e3), "unit": args.unit, "batch_per_gpu": 128,
                                    "ms_per_step": r32["ms"],
                                    "gemm": "fp32x3: reference-precision mode (1e-5), fp32 operands as bf16 (hi, lo) pairs, "
                                            "3 tcgen05 passes per tile, fp32 activations"}
            extras["sample_kl_alone"] = sample_kl_sweep(args, dev, pk)
            rb = measure(args, dev, 1, 0, local, batch=args.ref_batch, gemm="bf16", steps=10, warmup=3, timing=False, e2e=True)
            extras["reference_batch"] = {"batch_per_gpu": args.ref_batch, "value": args.ref_batch / (rb["ms"] / 1e3),
                                         "e2e": args.ref_batch / (rb["ms_e2e"] / 1e3), "unit": args.unit,
                                         "ms_per_step": rb["ms"],
                                         "cpu_reference_same_batch": None if cpu is None else cpu["value"]}
        except Exception as e:  # the headline line must still be printed
            extras["error"] = f"{type(e).__name__}: {e}"

    line = {"metric": args.metric, "value": value, "unit": args.unit, "n_gpus": world, "steps": args.steps,
            "warmup": res["n_warm"], "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.gemm == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args, world), "roofline": roofline, "roofline_sample_kl": roof_sk,
            "step_model_tflops": step_tf, "step_frac_of_gemm_roofline": step_tf / peak_tf / world, "kernels": kernels,
            "cpu_baseline": cpu,
            "e2e": None if e2e is None else {"value": e2e, "unit": args.unit, "ms_per_step": ms_e2e,
                                             "h2d_bytes_per_step": res["h2d_bytes"], "d2h_bytes_per_step": 12},
            "side_by_side": extras,
            "gpu_launches": res["launches"], "clocks": res["clocks"], "execution": res["graph_note"],
            "hbm_peak_gb": res["hbm_peak_gb"], "grad_allreduce_bytes_per_step": res["allreduce_bytes"],
            "grad_allreduce": {"launches_per_step": res["allreduce_launches"], "ms_alone": res["allreduce_alone_ms"],
                               "how": "the gradients that exist live in flat buckets, all-reduced after backward (NCCL, "
                                      "captured in the graph; not overlapped: a collective next to persistent "
                                      "one-CTA-per-SM contractions delays them by more than it saves); ms_alone = the "
                                      "same buckets all-reduced with nothing else running, CUDA events, after the timed "
                                      "region"}}
    print(json.dumps(line), flush=True)
    leave()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
