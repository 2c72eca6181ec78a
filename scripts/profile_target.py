"""Short, fixed workload for `ncu --set full`: the fused sample+KL kernel at
config-2 size and the three tcgen05 contractions at the BERT-base FFN shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bayeformers_b200 import _lib, ops
from bayeformers_b200._lib import BF_BF16, BF_PRIOR_GAUSSIAN, BF_PRIOR_MIXTURE
import numpy as np

lib = _lib.load()
REPS = int(os.environ.get("REPS", "2"))  # launches of each kernel (ncu --set full costs ~4 s per profiled launch)
DEV = "cuda:0"
st = torch.cuda.current_stream().cuda_stream
n, S = 4096 * 4096, 4
mu = torch.empty(n, device=DEV).uniform_(-0.2, 0.2)
rho = torch.empty(n, device=DEV).uniform_(-5, -4)
lq, lp = torch.empty(S, device=DEV), torch.empty(S, device=DEV)
moped = ops.PriorSpec(BF_PRIOR_GAUSSIAN, sigma1=1.3132616, mu=mu, rho=None)
mix = ops.PriorSpec(BF_PRIOR_MIXTURE, 0.5, 1.0, float(np.float32(np.exp(-6))))
for _ in range(REPS):
    ops.sample_kl_forward(mu, rho, moped, ops.StreamSpec(1, 2, 3), S, torch.bfloat16, lq, lp, False)
    ops.sample_kl_forward(mu, rho, mix, ops.StreamSpec(1, 2, 3), 1, torch.float32, lq, lp, False)
gw = torch.randn(S, n, device=DEV)
ops.sample_kl_backward(gw, mu, rho, ops.PriorSpec(), ops.StreamSpec(1, 2, 3), S, None, None, False)
del gw
S, M, N, K = 4, 4096, 3072, 768
x = torch.randn(S, M, K, device=DEV).bfloat16(); w = torch.randn(S, N, K, device=DEV).bfloat16()
gy = torch.randn(S, M, N, device=DEV).bfloat16()
y = torch.empty(S, M, N, device=DEV, dtype=torch.bfloat16); dx = torch.empty(S, M, K, device=DEV, dtype=torch.bfloat16)
mu2 = torch.randn(N, K, device=DEV) * 0.02; rho2 = torch.full((N, K), -5.0, device=DEV)
g_rho = torch.empty(N, K, device=DEV)
ws = torch.empty(lib.bf_linear_wgrad_fused_workspace_bytes(S, M, N, K, 0), dtype=torch.uint8, device=DEV)
for _ in range(REPS):
    lib.bf_linear_fwd(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), S, M, N, K, BF_BF16, BF_BF16, st)
    lib.bf_linear_dgrad(gy.data_ptr(), w.data_ptr(), dx.data_ptr(), S, M, N, K, BF_BF16, BF_BF16, st)
    lib.bf_linear_wgrad_fused(gy.data_ptr(), x.data_ptr(), S, M, N, K, BF_BF16, mu2.data_ptr(), rho2.data_ptr(), 2, None, None,
                              0.5, 1.0, 1.0, None, None, 7, 0, 1, None, None, g_rho.data_ptr(), 0, ws.data_ptr(), st)
# native LayerNorm at the BERT-base activation shape (S=4 x 64 sequences x 128 tokens)
rows, H = 4 * 64 * 128, 768
xa = torch.randn(rows, H, device=DEV).bfloat16().requires_grad_()
ga = torch.ones(4, H, device=DEV, requires_grad=True); ba = torch.zeros(4, H, device=DEV, requires_grad=True)
for _ in range(REPS):
    ya = ops.LayerNormFn.apply(xa, ga, ba, 4, 1e-12)
    ya.backward(torch.ones_like(ya))
# fused dropout + residual + LayerNorm (output blocks) at the same activation shape, p = 0.1, shared affine
ha = torch.randn(rows, H, device=DEV).bfloat16().requires_grad_()
ra = torch.randn(rows, H, device=DEV).bfloat16().requires_grad_()
gsh = torch.ones(H, device=DEV, requires_grad=True); bsh = torch.zeros(H, device=DEV, requires_grad=True)
for _ in range(REPS):
    yr = ops.ResidualLayerNormFn.apply(ha, ra, gsh, bsh, 4, 1e-12, ops.DropoutSpec(0.1, 1, 1, 1), [])
    yr.backward(torch.ones_like(yr))
# dgrad with the TMA reduce-add epilogue (gradient sinks) at the FFN shape
for _ in range(REPS):
    lib.bf_linear_dgrad_accumulate(gy.data_ptr(), w.data_ptr(), dx.data_ptr(), S, M, N, K, BF_BF16, BF_BF16, st)
# bias gradient at the FFN shape
db = torch.empty(S, N, device=DEV)
bws = torch.zeros(lib.bf_bias_grad_workspace_bytes(S, M, N), dtype=torch.uint8, device=DEV)
for _ in range(REPS):
    lib.bf_bias_grad(gy.data_ptr(), BF_BF16, db.data_ptr(), S, M, N, bws.data_ptr(), st)
# fused bias+GELU forward and GELU'/bias-grad backward at the FFN-up shape
z = torch.empty(S, M, N, device=DEV, dtype=torch.bfloat16); gz = torch.empty_like(z)
bias = torch.randn(S, N, device=DEV)
gws = torch.zeros(lib.bf_gelu_bwd_bias_grad_workspace_bytes(S, M, N), dtype=torch.uint8, device=DEV)
for _ in range(REPS):
    lib.bf_linear_fwd_gelu(x.data_ptr(), w.data_ptr(), bias.data_ptr(), z.data_ptr(), y.data_ptr(), S, M, N, K, st)
    lib.bf_gelu_bwd_bias_grad(gy.data_ptr(), z.data_ptr(), gz.data_ptr(), db.data_ptr(), S, M, N, gws.data_ptr(), st)
# fused clip + AdamW over 2 x 16.8 M fp32 parameters
import bayeformers_b200 as bf
pp = [torch.nn.Parameter(torch.randn(4096, 4096, device=DEV)) for _ in range(2)]
opt = bf.optim.ClipAdamW(pp, lr=1e-3, max_grad_norm=1.0)
for q in pp:
    q.grad = torch.randn_like(q)
for _ in range(REPS):
    opt.step()
# multi-tensor sample+KL over a 2 x (4096 x 4096 + bias) model, S = 4, MOPED prior, bf16 weights
net = torch.nn.Sequential(torch.nn.Linear(4096, 4096), torch.nn.Linear(4096, 4096))
bm = bf.to_bayesian(net, delta=0.05, freeze=True, gemm_dtype="bf16").to(DEV)
bf.enable_presample(bm)
for _ in range(REPS):
    bm._presampler.run(4)
torch.cuda.synchronize()
print("profile target done")
