"""Kernel-only timing (CUDA events per launch) of the multi-tensor sample+KL kernel on a 2 x (4096 x 4096 + bias)
model, S = 4, MOPED prior, bf16 weights (671 MB algorithmic per launch, > L2), and of the per-tensor fast kernel."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bayeformers_b200 as bf
from bayeformers_b200 import ops

DEV = "cuda:0"
PEAK = 6551.7
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", PEAK)
S = int(os.environ.get("S", 4))
net = torch.nn.Sequential(torch.nn.Linear(4096, 4096), torch.nn.Linear(4096, 4096))
bm = bf.to_bayesian(net, delta=0.05, freeze=True, gemm_dtype="bf16").to(DEV)
bf.enable_presample(bm)
for _ in range(3):
    bm._presampler.run(S)
torch.cuda.synchronize()
ops.enable_kernel_timing(True)
for _ in range(20):
    bm._presampler.run(S)
torch.cuda.synchronize()
k = ops.kernel_timing_summary()["sample_kl_fwd"]
ops.enable_kernel_timing(False)
ms = k["ms"] / k["calls"]
gbs = k["work"] / k["calls"] / ms / 1e6
print(f"multi  S={S} bf16 moped: {ms*1e3:7.1f} us  {gbs:7.1f} GB/s  frac {gbs/PEAK:.3f}   (BF_SK_PREFETCH={os.environ.get('BF_SK_PREFETCH','1')})")
if os.environ.get("MULTI_ONLY"):
    sys.exit(0)
# per-tensor fast kernel, same tensor size, constant-sigma MOPED prior and the default mixture prior
import numpy as np
from bayeformers_b200._lib import BF_PRIOR_GAUSSIAN, BF_PRIOR_MIXTURE
n = 4096 * 4096
mu = torch.empty(n, device=DEV).uniform_(-0.2, 0.2); rho = torch.empty(n, device=DEV).uniform_(-5, -4)
lq, lp = torch.empty(S, device=DEV), torch.empty(S, device=DEV)
for name, pr, P in (("moped", ops.PriorSpec(BF_PRIOR_GAUSSIAN, sigma1=1.3132616, mu=mu.clone(), rho=None), 4),
                    ("mixture", ops.PriorSpec(BF_PRIOR_MIXTURE, 0.5, 1.0, float(np.float32(np.exp(-6)))), 0)):
    for wd in (torch.bfloat16, torch.float32):
        for _ in range(3):
            ops.sample_kl_forward(mu, rho, pr, ops.StreamSpec(1, 2, 3), S, wd, lq, lp, False)
        torch.cuda.synchronize()
        ops.enable_kernel_timing(True)
        for _ in range(20):
            ops.sample_kl_forward(mu, rho, pr, ops.StreamSpec(1, 2, 3), S, wd, lq, lp, False)
        torch.cuda.synchronize()
        k = ops.kernel_timing_summary()["sample_kl_fwd"]
        ops.enable_kernel_timing(False)
        ms = k["ms"] / k["calls"]
        gbs = k["work"] / k["calls"] / ms / 1e6
        print(f"fast   S={S} {str(wd).split('.')[-1]:8s} {name:8s}: {ms*1e3:7.1f} us  {gbs:7.1f} GB/s  frac {gbs/PEAK:.3f}")
