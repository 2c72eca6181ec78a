"""Fixed workload for `ncu --set full` of the round-2 kernels AT THE SHAPES bench.py's default step launches them with
(BERT-base, B = 512 sequences x S = 4 samples x T = 128 tokens -> M = 65536 rows per sample): the fused-GELU forward and
the GELU' dgrad of the FFN layers, the plain forward of a 768 x 768 projection, the per-prior multi-tensor sample+KL
kernel, the row-sparse Embedding kernels at BERT-large size, the fp32x3 contraction and the native attention."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bayeformers_b200 as bf
from bayeformers_b200 import _lib, ops
from bayeformers_b200._lib import BF_BF16

lib = _lib.load()
DEV = "cuda:0"
st = torch.cuda.current_stream().cuda_stream
REPS = int(os.environ.get("REPS", "1"))
S, M, H, F = 4, 65536, 768, 3072
x = torch.randn(S, M, H, device=DEV).bfloat16()
w_up = (torch.randn(S, F, H, device=DEV) * 0.02).bfloat16()
w_dn = (torch.randn(S, H, F, device=DEV) * 0.02).bfloat16()
bias = torch.randn(S, F, device=DEV) * 0.02
z = torch.empty(S, M, F, device=DEV, dtype=torch.bfloat16)
y = torch.empty_like(z)
gy = torch.randn(S, M, H, device=DEV).bfloat16()
gz = torch.empty_like(z)
w_q = (torch.randn(S, H, H, device=DEV) * 0.02).bfloat16()
yq = torch.empty(S, M, H, device=DEV, dtype=torch.bfloat16)
for _ in range(REPS):
    lib.bf_linear_fwd_gelu(x.data_ptr(), w_up.data_ptr(), bias.data_ptr(), z.data_ptr(), y.data_ptr(), S, M, F, H, st)   # FFN-up fwd
    lib.bf_linear_dgrad_gelu(gy.data_ptr(), w_dn.data_ptr(), z.data_ptr(), gz.data_ptr(), S, M, H, F, st)             # FFN-down dgrad o gelu'
    lib.bf_linear_fwd(x.data_ptr(), w_q.data_ptr(), None, yq.data_ptr(), S, M, H, H, BF_BF16, BF_BF16, st)            # q / k / v fwd
torch.cuda.synchronize()
del z, y, gz
# fp32x3 (reference precision on the tensor cores) at a quarter of the rows
M3 = 16384
x3 = ops.split_bf16x2(torch.randn(S, M3, H, device=DEV))
w3 = ops.split_bf16x2(torch.randn(S, F, H, device=DEV) * 0.02)
y3 = torch.empty(S, M3, F, device=DEV)
for _ in range(REPS):
    lib.bf_linear_fwd_x3(x3[0].data_ptr(), x3[1].data_ptr(), w3[0].data_ptr(), w3[1].data_ptr(), None, y3.data_ptr(), S, M3, F, H, st)
del x3, w3, y3
# multi-tensor sample+KL (per-prior kernels): 2 x (4096 x 4096 + bias), S = 4, MOPED prior, bf16 samples
net = torch.nn.Sequential(torch.nn.Linear(4096, 4096), torch.nn.Linear(4096, 4096))
bm = bf.to_bayesian(net, delta=0.05, freeze=True, gemm_dtype="bf16").to(DEV)
bf.enable_presample(bm)
for _ in range(REPS + 1):
    bm._presampler.run(4)
# row-sparse Bayesian Embedding at BERT-large size: V = 30522, H = 1024, 65536 tokens of S = 16 samples
import bayeformers_b200.nn as bnn
emb = bnn.Embedding.from_frequentist(torch.nn.Embedding(30522, 1024, padding_idx=0), delta=0.05, freeze=True).to(DEV)
emb.gemm_dtype = torch.bfloat16
ids = torch.randint(0, 30522, (16 * 8, 512), device=DEV)
for _ in range(REPS):
    with bf.mc_samples(16):
        e = emb(ids)
    e.backward(torch.ones_like(e))
    emb.zero_grad(set_to_none=True)
# native attention at the bench shape: 2048 folded sequences x 12 heads x 128 tokens
B2 = 2048
qkv = [(torch.randn(B2, 128, 768, device=DEV) * 0.5).bfloat16().requires_grad_() for _ in range(3)]
q, k, v = (t.view(B2, 128, 12, 64).transpose(1, 2) for t in qkv)
for _ in range(REPS):
    o = ops.AttentionFn.apply(q, k, v, 0.125, ops.DropoutSpec(0.1, 1, 2, 3))
    o.backward(torch.ones_like(o))
torch.cuda.synchronize()
print("profile_r02 done")
