"""Fixed workload for `ncu --set full` of the attention kernels at the bench shape (2048 folded sequences x 12 heads x
128 tokens x 64, dropout 0.1): the tcgen05 forward / backward (default), the mma.sync kernels (BF_OPT_ATTN_TC = 0) and
torch's fused SDPA (cuDNN) forward / backward for comparison."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from bayeformers_b200 import _lib, ops

lib = _lib.load()
DEV = "cuda:0"
B, T, H, Dh = 2048, 128, 12, 64
torch.manual_seed(0)
qkv = [(torch.randn(B, T, H * Dh, device=DEV) * 0.5).bfloat16().requires_grad_() for _ in range(3)]
q, k, v = (t.view(B, T, H, Dh).transpose(1, 2) for t in qkv)
gout = torch.randn(B, T, H, Dh, device=DEV).bfloat16()
for tc in (1, 0):
    lib.bf_set_option(_lib.BF_OPT_ATTN_TC, tc)
    for _ in range(2):
        o = ops.AttentionFn.apply(q, k, v, 0.125, ops.DropoutSpec(0.1, 1, 2, 3))
        o.backward(gout)
lib.bf_set_option(_lib.BF_OPT_ATTN_TC, 1)
for _ in range(2):
    o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.1)
    o.backward(gout.transpose(1, 2))
torch.cuda.synchronize()
print("profile_attention done")
