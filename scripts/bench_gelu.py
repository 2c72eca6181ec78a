"""Device times (CUDA events, median of 10) of the FFN contractions at the bench shape (S=4, M=65536, 768 <-> 3072):
plain forward, forward + GELU epilogue, plain dgrad, dgrad + GELU' epilogue."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bayeformers_b200 import _lib
from bayeformers_b200._lib import BF_BF16

lib = _lib.load()
DEV = "cuda:0"
st = torch.cuda.current_stream().cuda_stream
S, M, H, F = 4, 65536, 768, 3072
x = torch.randn(S, M, H, device=DEV).bfloat16()
w_up = (torch.randn(S, F, H, device=DEV) * 0.02).bfloat16()
w_dn = (torch.randn(S, H, F, device=DEV) * 0.02).bfloat16()
bias = torch.randn(S, F, device=DEV) * 0.02
z = torch.empty(S, M, F, device=DEV, dtype=torch.bfloat16)
y = torch.empty_like(z)
gy = torch.randn(S, M, H, device=DEV).bfloat16()
gz = torch.empty_like(z)


def timed(fn, reps=10):
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


out = {}
lib.bf_set_option(_lib.BF_OPT_GELU_POLY, 0)
out["fwd_gelu_erf_ms"] = timed(lambda: lib.bf_linear_fwd_gelu(x.data_ptr(), w_up.data_ptr(), bias.data_ptr(), z.data_ptr(), y.data_ptr(), S, M, F, H, st))
out["dgrad_gelu_erf_ms"] = timed(lambda: lib.bf_linear_dgrad_gelu(gy.data_ptr(), w_dn.data_ptr(), z.data_ptr(), gz.data_ptr(), S, M, H, F, st))
lib.bf_set_option(_lib.BF_OPT_GELU_POLY, 1)
out["fwd_plain_ms"] = timed(lambda: lib.bf_linear_fwd(x.data_ptr(), w_up.data_ptr(), bias.data_ptr(), z.data_ptr(), S, M, F, H, BF_BF16, BF_BF16, st))
out["fwd_gelu_ms"] = timed(lambda: lib.bf_linear_fwd_gelu(x.data_ptr(), w_up.data_ptr(), bias.data_ptr(), z.data_ptr(), y.data_ptr(), S, M, F, H, st))
out["dgrad_plain_ms"] = timed(lambda: lib.bf_linear_dgrad(gy.data_ptr(), w_dn.data_ptr(), gz.data_ptr(), S, M, H, F, BF_BF16, BF_BF16, st))
out["dgrad_gelu_ms"] = timed(lambda: lib.bf_linear_dgrad_gelu(gy.data_ptr(), w_dn.data_ptr(), z.data_ptr(), gz.data_ptr(), S, M, H, F, st))
flop = 2.0 * S * M * H * F
out["tflops"] = {k[:-3]: flop / (v * 1e9) for k, v in out.items() if k.endswith("_ms")}
print(json.dumps(out))
