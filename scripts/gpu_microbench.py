"""Device-timed micro-benchmarks of the sample+KL kernels (CUDA events, L2 flushed
by input size: 16.8M elements x >= 12 B > 126 MB L2)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bayeformers_b200 import ops
from bayeformers_b200._lib import BF_PRIOR_GAUSSIAN, BF_PRIOR_MIXTURE, BF_PRIOR_NONE

DEV = "cuda:0"
PEAK = 6551.7
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", PEAK)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    n = 4096 * 4096 + 4096
    mu = torch.empty(n, device=DEV).uniform_(-0.2, 0.2)
    rho = torch.empty(n, device=DEV).uniform_(-5, -4)
    pmu = mu.clone()
    prho = torch.ones(n, device=DEV)
    rows = []
    for prior_name in ("mixture", "gaussian"):
        for S in (1, 2, 4, 8, 16):
            for wd in (torch.float32, torch.bfloat16):
                if prior_name == "mixture":
                    pr = ops.PriorSpec(BF_PRIOR_MIXTURE, 0.5, 1.0, float(np.float32(np.exp(-6))))
                    P = 0
                else:
                    pr = ops.PriorSpec(BF_PRIOR_GAUSSIAN, mu=pmu, rho=prho)
                    P = 8
                logq, logp = torch.empty(S, device=DEV), torch.empty(S, device=DEV)
                st = ops.StreamSpec(1, 2, 3)
                bw = 2 if wd == torch.bfloat16 else 4
                w = torch.empty(S, n, dtype=wd, device=DEV)

                def f():
                    ops.sample_kl_forward(mu, rho, pr, st, S, wd, logq, logp, False)

                ms = timeit(f)
                bytes_ = n * (8 + P + S * bw) + 8 * S
                gbs = bytes_ / ms / 1e6
                rows.append(dict(kernel="sample_kl_fwd", prior=prior_name, S=S, w=str(wd).split(".")[-1], ms=round(ms, 4),
                                 GBs=round(gbs, 1), frac=round(gbs / PEAK, 3), Gelem_s=round(n * S / ms / 1e6, 1)))
                print(rows[-1], flush=True)
    for S in (1, 4):
        for gd in (torch.float32, torch.bfloat16):
            gw = torch.randn(S, n, device=DEV).to(gd)
            st = ops.StreamSpec(1, 2, 3)

            def f():
                ops.sample_kl_backward(gw, mu, rho, ops.PriorSpec(), st, S, None, None, False)

            ms = timeit(f)
            bytes_ = n * (S * gw.element_size() + 4 + 4)
            gbs = bytes_ / ms / 1e6
            rows.append(dict(kernel="sample_kl_bwd", S=S, gw=str(gd).split(".")[-1], ms=round(ms, 4), GBs=round(gbs, 1),
                             frac=round(gbs / PEAK, 3)))
            print(rows[-1], flush=True)
    # plain copy for calibration on this box
    a = torch.empty(1 << 28, device=DEV); b = torch.empty_like(a)
    ms = timeit(lambda: b.copy_(a))
    print(dict(kernel="torch_copy_1GiB", ms=round(ms, 4), GBs=round(2 * a.numel() * 4 / ms / 1e6, 1)))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/microbench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
