#!/usr/bin/env python
"""bench.py -- headline benchmark of the variational-layer hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Workload (BASELINE.json configs[2], the one the metric is quoted on): Bayesian
BERT-base sequence classification, `to_bayesian(delta=0.05, freeze=True)`, synthetic
tokens of length 128, S=4 Monte-Carlo samples, training step = S-sample forward
+ ELBO loss + backward + grad-clip + AdamW (the pattern of
/root/reference/examples/bert_glue.py:56-73,225-241).  One JSON line on stdout.
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

METRIC = "bayes_bert_base_train_seqs_per_s"
N_BATCHES = 1000  # the reference divides the KL term by len(train_loader) (bert_glue.py:235)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="sequences per GPU per step (before the S-fold); 512 -> 89 GB of HBM")
    ap.add_argument("--graph", type=int, default=1,
                    help="1: capture the whole training step in one CUDA graph; 0: eager")
    ap.add_argument("--samples", type=int, default=4)
    ap.add_argument("--seq", type=int, default=128)
    ap.add_argument("--gemm", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--kl-grad", type=int, default=1)
    ap.add_argument("--ref-batch", type=int, default=2, help="sequences per step of the CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fused-optim", type=int, default=1, help="bf.optim.ClipAdamW instead of clip_grad_norm_ + AdamW")
    ap.add_argument("--presample", type=int, default=1, help="one multi-tensor sample+KL launch per forward")
    ap.add_argument("--fuse-gelu", type=int, default=1,
                    help="move the FFN GELU into the Bayesian Linear (fused tensor-core epilogue)")
    ap.add_argument("--host-ln", type=int, default=1,
                    help="route the host model's frequentist LayerNorms through the native LayerNorm kernels")
    ap.add_argument("--fuse-residual", type=int, default=1,
                    help="fuse dropout + residual add + LayerNorm (+ the Linear's bias gradient) of the HF output blocks")
    ap.add_argument("--grad-sinks", type=int, default=1,
                    help="accumulate the Linear dgrads of a residual-shared input in place (TMA reduce-add) instead of "
                         "autograd's separate add passes; needs --fuse-residual 1")
    ap.add_argument("--layers", type=int, default=0, help="debug: override num_hidden_layers")
    ap.add_argument("--profile", action="store_true",
                    help="for runs under ncu only: allow < 3 warm-up steps, skip the e2e and CPU legs (numbers invalid)")
    return ap.parse_args()


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


def build_bert(layers: int = 0):
    from transformers import BertConfig, BertForSequenceClassification

    cfg = BertConfig(num_labels=2)
    if layers:
        cfg.num_hidden_layers = layers
    torch.manual_seed(0)
    model = BertForSequenceClassification(cfg)
    with torch.no_grad():  # HF zero-inits biases; perturb so MOPED sees non-degenerate values
        g = torch.Generator().manual_seed(1)
        for n, p in model.named_parameters():
            if n.endswith("bias"):
                p.add_(torch.randn(p.shape, generator=g) * 0.02)
    return model, cfg


def flops_per_seq_sample(cfg, T):
    """fwd+bwd matmul flops of one sequence for ONE MC sample: 6*T*params_linear + attention."""
    H, L, FF = cfg.hidden_size, cfg.num_hidden_layers, cfg.intermediate_size
    lin = L * (4 * H * H + 2 * H * FF) + H * H + H * cfg.num_labels
    return 3 * (2 * T * lin), 3 * (L * 4 * T * T * H)


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
def cpu_reference_steps(args, steps: int, warmup: int):
    """The reference's CPU path for this workload through the oracle port
    (oracle/bayes_oracle.py: same torch-CPU operator sequence as
    bayeformers/nn/layers/linear.py:83-104 inside the S-loop of
    examples/bert_glue.py:56-73), on all host cores, on a bounded sample of
    `--ref-batch` sequences per step.  Returns (seq_per_s, s_per_step, cores)."""
    from oracle import bayes_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, cfg = build_bert(args.layers)
    torch.manual_seed(2)
    om = O.oracle_convert(model, delta=0.05, freeze=True).train()
    B, T, S = args.ref_batch, args.seq, args.samples
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, cfg.vocab_size, (B, T), generator=g)
    labels = torch.randint(0, 2, (B,), generator=g)
    params = [p for p in om.parameters() if p.requires_grad]
    optim = torch.optim.AdamW(params, lr=2e-5, eps=1e-8)

    def step():  # same work as the GPU step: S-loop fwd, ELBO, bwd, clip, AdamW (bert_glue.py:230-241)
        optim.zero_grad(set_to_none=True)
        O.s_loop_step(om, lambda m: m(input_ids=ids).logits, labels, S, N_BATCHES)
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        optim.step()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return B / dt, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    seqs, dt, cores = cpu_reference_steps(args, args.steps, args.warmup)
    sample = (f"{args.ref_batch} sequences x S={args.samples} per step (bounded sample of the per-GPU batch), "
              f"oracle port of the reference S-loop, torch CPU fp32, {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": seqs, "unit": "seq/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": seqs, "unit": "seq/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": seqs, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": "BERT-base to_bayesian(delta=0.05, freeze=True) GLUE-style classification, synthetic tokens, "
                        "training step (S-sample fwd + ELBO + bwd + clip + AdamW)",
            "seq_len": args.seq, "mc_samples": args.samples, "batch_per_gpu": args.batch,
            "global_batch": args.batch * args.gpus, "gemm": args.gemm, "kl_grad": bool(args.kl_grad),
            "optimizer": "bf.optim.ClipAdamW (fused clip + AdamW)" if args.fused_optim else "clip_grad_norm_ + torch AdamW(fused)",
            "sampling": "multi-tensor (1 launch per forward)" if args.presample else "per layer",
            "ffn_gelu": "fused into bnn.Linear (epilogue + GELU'/bias-grad pass)" if args.fuse_gelu else "torch",
            "host_layernorm": "native kernels (bf_layernorm_*)" if args.host_ln else "torch",
            "output_blocks": "dropout + residual + LayerNorm fused (bf_resln_*, Philox mask, bias grad handed to the Linear)"
                             if args.fuse_residual else "torch dropout + add, separate LayerNorm",
            "shared_input_grads": "accumulated in place by the dgrad kernels (TMA reduce-add)"
                                  if (args.grad_sinks and args.fuse_residual) else "autograd add passes",
            "parallelism": f"dp{args.gpus} (batch sharded, identical Philox weights per rank, NCCL grad all-reduce)",
            "l2": "working set (0.7 GB sampled weights + tens of GB of activations) far exceeds the 126 MB L2; no explicit flush"}


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist

    import bayeformers_b200 as bf
    from bayeformers_b200 import ops, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    if os.environ.get("BF_ANOMALY"):  # debugging aid: forward traceback of a failing backward node
        torch.autograd.set_detect_anomaly(True)

    model, cfg = build_bert(args.layers)
    bf.manual_seed(1234)
    bm = bf.to_bayesian(model, delta=0.05, freeze=True, gemm_dtype=args.gemm, kl_grad=bool(args.kl_grad))
    if args.host_ln or args.fuse_gelu or args.fuse_residual:
        # same parameters and numerics: native LayerNorm kernels (fp32 gamma/beta); FFN GELU fused into the layer;
        # dropout + residual + LayerNorm of the output blocks in one pass each way (Philox dropout mask)
        bf.accelerate_host_(bm, layernorm=bool(args.host_ln), fuse_gelu=bool(args.fuse_gelu),
                            fuse_residual=bool(args.fuse_residual),
                            grad_sinks=bool(args.grad_sinks and args.fuse_residual))
    bm = bm.to(dev).train()
    if args.presample:
        bf.enable_presample(bm)
    if world > 1:
        parallel.broadcast_seed(0)
    if args.gemm == "bf16":
        # activations flow in bf16 (embeddings / LayerNorm of the host model cast to bf16);
        # the variational masters (mu, rho, priors) stay fp32
        bf.cast_frequentist_(bm, torch.bfloat16)
    params = [p for p in bm.parameters() if p.requires_grad]
    use_graph = bool(args.graph) and not args.profile
    if args.fused_optim:  # global-norm clip + AdamW in two launches (section 8f row 3)
        optim = bf.optim.ClipAdamW(params, lr=2e-5, eps=1e-8, weight_decay=0.01, max_grad_norm=1.0)
    else:
        optim = torch.optim.AdamW(params, lr=2e-5, eps=1e-8, fused=True, capturable=use_graph)
    sync = parallel.GradSync(bm)
    bf.enable_device_step(dev)  # eps = f(seed, tensor, host_step + device_step, sample): graph replays draw fresh eps

    B, T, S = args.batch, args.seq, args.samples
    g = torch.Generator().manual_seed(100 + rank)
    ids_host = torch.randint(0, cfg.vocab_size, (B, T), generator=g).pin_memory()
    labels_host = torch.randint(0, 2, (B,), generator=g).pin_memory()
    ids_dev, labels_dev = ids_host.to(dev), labels_host.to(dev)
    out_host = torch.zeros(3, dtype=torch.float32).pin_memory()

    def step_body(ids, labels):
        bf.advance_step()
        optim.zero_grad(set_to_none=True)
        with bf.mc_samples(S):
            logits = bm(input_ids=ids.repeat(S, 1)).logits
        raw = logits.float().view(S, B, -1)
        nll = torch.nn.functional.cross_entropy(raw.mean(0), labels)
        lp, lq = bm.log_prior().mean(), bm.log_variational_posterior().mean()
        loss = (lq - lp) / N_BATCHES + nll
        loss.backward()
        sync.finish()
        if not args.fused_optim:
            torch.nn.utils.clip_grad_norm_(params, 1.0)
        optim.step()
        return loss, lp, lq

    graph_note = "eager"
    step = step_body
    if use_graph:
        try:
            static_ids, static_labels = ids_dev.clone(), labels_dev.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                # every Linear contraction is ours, so nothing has touched cuBLAS yet; SDPA may fall back to
                # bmm under capture, and creating a cuBLAS handle while capturing is illegal: make it exist now
                _d = torch.ones(64, 64, device=dev, dtype=torch.bfloat16)
                torch.bmm(_d[None], _d[None]); torch.mm(_d.float(), _d.float())
                for _ in range(3):
                    step_body(static_ids, static_labels)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            optim.zero_grad(set_to_none=True)
            l0 = ops.stats["launches"]
            # bf.hf_capture_compat: keep HF on the mask-free fused-attention path while capturing (see its docstring)
            with bf.hf_capture_compat(), torch.cuda.graph(graph, stream=side):  # same stream as the warm-up
                static_out = step_body(static_ids, static_labels)
            launches_per_graph = ops.stats["launches"] - l0

            def step(ids, labels):  # noqa: F811
                if ids is not static_ids:
                    static_ids.copy_(ids, non_blocking=True)
                    static_labels.copy_(labels, non_blocking=True)
                graph.replay()
                ops.stats["launches"] += launches_per_graph
                return static_out

            ids_dev, labels_dev = static_ids, static_labels
            graph_note = "whole training step captured in one CUDA graph"
        except Exception as e:  # e.g. a host sync inside the host model: restart this process in eager mode
            import traceback
            traceback.print_exc(file=sys.stderr)
            sys.stderr.write(f"[bench] CUDA-graph capture failed ({type(e).__name__}); re-running eagerly\n")
            sys.stderr.flush()
            if world == 1:
                os.execv(sys.executable, [sys.executable] + sys.argv + ["--graph", "0"])
            use_graph, step, graph_note = False, step_body, "eager (graph capture failed)"
            ids_dev, labels_dev = ids_host.to(dev), labels_host.to(dev)
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_warm = args.warmup if args.profile else max(args.warmup, 3)
    for _ in range(n_warm):
        step(ids_dev, labels_dev)

    # ---- timed region 1: inputs resident in HBM (value)
    ops.enable_kernel_timing(not use_graph)
    launches0 = ops.stats["launches"]
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profile:  # `ncu --profile-from-start off`: only the measured step(s) are profiled
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        step(ids_dev, labels_dev)
    e1.record()
    barrier()
    if args.profile:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    clk = clocks.stop() if clocks else None
    launches = ops.stats["launches"] - launches0
    kern = ops.kernel_timing_summary()
    ops.enable_kernel_timing(False)
    kern_steps = args.steps
    if use_graph:
        # CUDA events cannot bracket kernels inside a replayed graph: take the per-kernel durations from
        # eager, instrumented executions of the same step right after the timed region
        kern_steps = 2
        step_body(ids_dev, labels_dev)
        barrier()
        ops.enable_kernel_timing(True)
        for _ in range(kern_steps):
            step_body(ids_dev, labels_dev)
        barrier()
        kern = ops.kernel_timing_summary()
        ops.enable_kernel_timing(False)

    # ---- timed region 2: end to end through the public API with HOST buffers (e2e)
    barrier()
    e0.record()
    for _ in range(0 if args.profile else args.steps):
        ids = ids_host.to(dev, non_blocking=True)
        labels = labels_host.to(dev, non_blocking=True)
        loss, lp, lq = step(ids, labels)
        out_host.copy_(torch.stack([loss.detach(), lp.detach().float(), lq.detach().float()]), non_blocking=True)
        torch.cuda.synchronize()
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 1e-9) / args.steps

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    def leave():
        """End of a multi-rank run.  Tearing NCCL down while a CUDA graph that captured its collectives is still
        alive can hang at exit: drop the graph first, and leave without running destructors."""
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

    value = world * B / (ms / 1e3)
    e2e = world * B / (ms_e2e / 1e3)
    # ---- roofline of the dominant kernel family (tcgen05 contractions), from the CUDA-event brackets
    gemm = {k: v for k, v in kern.items() if k.startswith("gemm_")}
    g_ms = sum(v["ms"] for v in gemm.values())
    g_flops = sum(v["work"] for v in gemm.values())
    g_calls = sum(v["calls"] for v in gemm.values())
    peak_tf = pk["bf16_tflops_sustained"]
    ach_tf = g_flops / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "tcgen05 contractions: tc2::bayes_gemm2_kernel (fwd, dgrad; cta_group::2) + wg::bayes_wgrad_kernel (fused wgrad), all layers",
                "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                "peak_source": f"{pk['source']} bf16_tflops_sustained (kernel timed inside a long step; the peak is cuBLAS "
                               "running back to back for 4 s under the power cap, these launches are interleaved with "
                               "lighter kernels and can clock higher, so frac may slightly exceed 1)",
                # dram__bytes_read+write of the profiled FFN-shape fwd launch (S=4, M=4096, N=3072, K=768; algorithmic
                # operand+result bytes 145 MB, most of the bf16 result still in L2 at kernel end):
                # profiles/r01e_ncu_full_kernels_final.md row 3 (44.1 + 48.7 MB)
                "traffic": 92.8e6, "launches_per_step": g_calls / kern_steps, "avg_launch_ms": g_ms / max(g_calls, 1),
                "share_of_step": g_ms / kern_steps / ms,
                "timing": "CUDA events around every launch on the launching stream" +
                          (", taken in eager executions of the same step after the timed graph replays" if use_graph else
                           ", inside the timed region")}
    sk = kern.get("sample_kl_fwd")
    roof_sk = None
    if sk and sk["ms"] > 0:
        gbs = sk["work"] / (sk["ms"] / 1e3) / 1e9
        roof_sk = {"bound": "hbm", "kernel": "sample_kl_multi_kernel" if args.presample else "sample_kl_fwd_fast_kernel", "achieved": gbs, "peak": pk["hbm_gbs"],
                   "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "share_of_step": sk["ms"] / kern_steps / ms,
                   "traffic": None}
    lin_f, att_f = flops_per_seq_sample(cfg, T)
    step_tf = (lin_f + att_f) * S * value / 1e12
    kernels = {k: {"calls_per_step": v["calls"] / kern_steps, "ms_per_step": v["ms"] / kern_steps,
                   "share": v["ms"] / kern_steps / ms} for k, v in sorted(kern.items())}

    cpu = None
    if world == 1 and not args.no_cpu_baseline and not args.profile:
        seqs, dt, cores = cpu_reference_steps(args, 2, 1)
        cpu = {"value": seqs, "unit": "seq/s", "cores": cores, "kind": "port",
               "sample": f"{args.ref_batch} sequences x S={args.samples}, 1 warm-up + 2 timed steps of the oracle port "
                         f"of the reference S-loop ({dt:.2f} s/step), torch CPU fp32"}

    line = {"metric": METRIC, "value": value, "unit": "seq/s", "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.gemm == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args), "roofline": roofline, "roofline_sample_kl": roof_sk,
            "step_model_tflops": step_tf, "step_frac_of_gemm_roofline": step_tf / peak_tf, "kernels": kernels,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": "seq/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": ids_host.numel() * 8 + labels_host.numel() * 8, "d2h_bytes_per_step": 12},
            "gpu_launches": launches, "clocks": clk, "execution": graph_note,
            "hbm_peak_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
            "grad_allreduce_bytes_per_step": sync.bytes_last_step}
    print(json.dumps(line), flush=True)
    leave()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
