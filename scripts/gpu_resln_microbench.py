"""Stand-alone timing of the fused dropout + residual + LayerNorm kernels at the bench shape
(S=4, rows = 4*256*128, H=768, bf16, p=0.1) against the torch composition they replace.
CUDA events, 3 warm-up + 20 timed launches, inputs (805 MB per pass set) far larger than L2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bayeformers_b200 as bf
from bayeformers_b200 import ops

dev = "cuda:0"
S, M, H = 4, int(os.environ.get("M", 32768)), int(os.environ.get("H", 768))
dt = torch.bfloat16
h = torch.randn(S * M, H, device=dev, dtype=dt).requires_grad_()
r = torch.randn(S * M, H, device=dev, dtype=dt).requires_grad_()
g = torch.ones(H, device=dev).requires_grad_()
b = torch.zeros(H, device=dev).requires_grad_()
gy = torch.randn(S * M, H, device=dev, dtype=dt)
spec = ops.DropoutSpec(0.1, 1, 1, 1)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


pass_bytes = S * M * H * 2
box = []
# kernel-only durations: CUDA events around each C-ABI launch (ops._timed), not around autograd's bookkeeping
# (for leaf inputs backward() also runs AccumulateGrad adds, which are not part of the kernel)
hn, rn = h.detach(), r.detach().requires_grad_()  # one input must need grad for backward to run
for _ in range(3):
    ops.ResidualLayerNormFn.apply(hn, rn, g, b, S, 1e-12, spec, box).backward(gy)
    rn.grad = None
torch.cuda.synchronize()
ops.enable_kernel_timing(True)
for _ in range(20):
    ops.ResidualLayerNormFn.apply(hn, rn, g, b, S, 1e-12, spec, box).backward(gy)
    rn.grad = None
torch.cuda.synchronize()
k = ops.kernel_timing_summary()
ops.enable_kernel_timing(False)
t_f, t_b = k["resln_fwd"]["ms"] / k["resln_fwd"]["calls"], k["resln_bwd"]["ms"] / k["resln_bwd"]["calls"]
print(f"fused   fwd {t_f*1e3:8.1f} us  ({4*pass_bytes/t_f/1e6:7.1f} GB/s on 4 passes)   "
      f"bwd {t_b*1e3:8.1f} us ({4*pass_bytes/t_b/1e6:7.1f} GB/s on 4 passes)   [kernel-only, CUDA events per launch]")
if os.environ.get("FUSED_ONLY"):
    sys.exit(0)
ln = torch.nn.LayerNorm(H, eps=1e-12).to(dev).to(dt)
hl = bf.accelerate_host_(torch.nn.Sequential(torch.nn.LayerNorm(H, eps=1e-12))).to(dev)[0]
for name, mod in (("torch LN", ln), ("native LN", hl)):
    f = lambda: mod(torch.nn.functional.dropout(h, 0.1, True) + r)
    t_f = timeit(f)
    t_fb = timeit(lambda: f().backward(gy))
    print(f"unfused ({name:9s}) fwd {t_f*1e3:8.1f} us   bwd {(t_fb-t_f)*1e3:8.1f} us  (+ bias_grad pass not included)")
