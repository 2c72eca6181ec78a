import sys, json
sys.path.insert(0, ".")
import torch
from bayeformers_b200 import _lib
lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
S, M, H, F = 4, 65536, 768, 3072
x = torch.randn(S, M, H, device="cuda").bfloat16(); w_up = (torch.randn(S, F, H, device="cuda") * 0.02).bfloat16()
w_dn = (torch.randn(S, H, F, device="cuda") * 0.02).bfloat16(); bias = torch.randn(S, F, device="cuda") * 0.02
z = torch.empty(S, M, F, device="cuda", dtype=torch.bfloat16); y = torch.empty_like(z)
gy = torch.randn(S, M, H, device="cuda").bfloat16(); gz = torch.empty_like(z)
def run(fn, reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
fwd = lambda: lib.bf_linear_fwd_gelu(x.data_ptr(), w_up.data_ptr(), bias.data_ptr(), z.data_ptr(), y.data_ptr(), S, M, F, H, st)
dg = lambda: lib.bf_linear_dgrad_gelu(gy.data_ptr(), w_dn.data_ptr(), z.data_ptr(), gz.data_ptr(), S, M, H, F, st)
for _ in range(30): fwd(); dg()
res = {"fwd": {0: [], 1: []}, "dgrad": {0: [], 1: []}}
for rnd in range(6):
    for poly in (0, 1):
        lib.bf_set_option(_lib.BF_OPT_GELU_POLY, poly)
        res["fwd"][poly].append(run(fwd)); res["dgrad"][poly].append(run(dg))
for k, v in res.items():
    print(k, "erf", ["%.3f" % t for t in v[0]], "poly", ["%.3f" % t for t in v[1]])
