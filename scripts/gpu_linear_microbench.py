"""BASELINE.json configs[1]: bnn.Linear 4096x4096, batch 8192, MC-sample sweep S = 1..32, fwd+bwd.
Default init (Uniform) and default scale-mixture prior, x ~ N(0,1).  Reports ms per fwd+bwd, contraction
TFLOP/s (6*M*N*K per sample) and the sample+KL kernel's GB/s (CUDA events around each launch)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bayeformers_b200 as bf
import bayeformers_b200.nn as bnn
from bayeformers_b200 import ops

DEV = "cuda:0"
torch.manual_seed(0); bf.manual_seed(1)
N = K = 4096; B = 8192
out = []
for mode in ("bf16", "fp32"):
    for S in ((1, 2, 4, 8, 16, 32) if mode == "bf16" else (1, 4)):
        layer = bnn.Linear(K, N).to(DEV)
        layer.gemm_dtype = bf.runtime._as_dtype(mode); layer.kl_grad = True
        x = torch.randn(S * B, K, device=DEV, dtype=torch.bfloat16 if mode == "bf16" else torch.float32).requires_grad_()
        def step():
            with bf.mc_samples(S):
                y = layer(x)
            loss = y.float().square().mean() + 1e-6 * (layer.live_log_variational_posterior - layer.live_log_prior).mean()
            loss.backward()
            layer.zero_grad(set_to_none=True); x.grad = None
        for _ in range(2): step()
        torch.cuda.synchronize()
        it = 5 if S <= 8 else 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it): step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        ops.enable_kernel_timing(True); step(); torch.cuda.synchronize()
        k = ops.kernel_timing_summary(); ops.enable_kernel_timing(False)
        sk = k.get("sample_kl_fwd", {"ms": 0, "work": 0})
        g_ms = sum(v["ms"] for n, v in k.items() if n.startswith("gemm_"))
        g_fl = sum(v["work"] for n, v in k.items() if n.startswith("gemm_"))
        row = {"gemm": mode, "S": S, "ms_fwd_bwd": round(ms, 3), "step_TFLOPs": round(6.0 * S * B * N * K / ms / 1e9, 1),
               "contractions_TFLOPs": round(g_fl / max(g_ms, 1e-9) / 1e9, 1),
               "sample_kl_ms": round(sk["ms"], 4), "sample_kl_GBs": round(sk["work"] / max(sk["ms"], 1e-9) / 1e6, 1),
               "kernels_ms": {n: round(v["ms"], 3) for n, v in sorted(k.items())}}
        print(json.dumps(row), flush=True); out.append(row)
        del layer, x
        torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/linear_microbench.json", "w"), indent=1)
