"""Device times of the native attention kernels against torch's fused SDPA (cuDNN / flash) at the bench shape:
2048 folded sequences x 12 heads x 128 tokens x 64, dropout 0.1, bf16.  CUDA events, L2 flushed between launches by the
size of the tensors themselves (q, k, v, o: 4 x 403 MB)."""
import json
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from bayeformers_b200 import ops  # noqa: E402

DEV = torch.device("cuda:0")
B, T, H, Dh = 2048, 128, 12, 64
REPS = 10


def timed(fn, reps=REPS):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]


def main():
    torch.manual_seed(0)
    qkv = [(torch.randn(B, T, H * Dh, device=DEV) * 0.5).bfloat16().requires_grad_() for _ in range(3)]
    q, k, v = (t.view(B, T, H, Dh).transpose(1, 2) for t in qkv)
    gout = torch.randn(B, T, H, Dh, device=DEV).bfloat16()
    out = {}
    drop = ops.DropoutSpec(0.1, 1, 2, 3)
    with torch.no_grad():
        out["native_fwd_ms"] = timed(lambda: ops.AttentionFn.apply(q, k, v, 0.125, drop))

    def native():
        o = ops.AttentionFn.apply(q, k, v, 0.125, drop)
        o.backward(gout)
    out["native_fwd_bwd_ms"] = timed(native)

    # the kernels alone (CUDA events around each C-ABI call)
    for t in qkv:
        t.grad = None
    ops.enable_kernel_timing(True)
    for _ in range(REPS):
        native()
    torch.cuda.synchronize()
    for name, d in ops.kernel_timing_summary().items():
        out[name + "_kernel_ms"] = d["ms"] / d["calls"]
    ops.enable_kernel_timing(False)

    gout_t = gout.transpose(1, 2)
    with torch.no_grad():
        out["sdpa_fwd_ms"] = timed(lambda: F.scaled_dot_product_attention(q, k, v, dropout_p=0.1))

    def sdpa():
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.1)
        o.backward(gout_t)
    out["sdpa_fwd_bwd_ms"] = timed(sdpa)
    gb = B * T * H * Dh * 2 / 1e9
    out["hbm_floor_ms"] = {"fwd": 4 * gb / 6.5517, "bwd": 7 * gb / 6.5517}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
