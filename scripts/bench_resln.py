import sys; sys.path.insert(0, ".")
import torch
from bayeformers_b200 import _lib
from bayeformers_b200._lib import BF_BF16
lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
S, M, H = 4, 65536, 768
rows = S * M
h = torch.randn(rows, H, device="cuda").bfloat16(); r = torch.randn(rows, H, device="cuda").bfloat16()
g = torch.ones(H, device="cuda"); b = torch.zeros(H, device="cuda")
z = torch.empty_like(h); y = torch.empty_like(h); mean = torch.empty(rows, device="cuda"); rstd = torch.empty(rows, device="cuda")
keep = torch.empty(rows, 32, dtype=torch.int32, device="cuda")
gy = torch.randn(rows, H, device="cuda").bfloat16(); dz = torch.empty_like(h); dh = torch.empty_like(h)
dg = torch.empty(H, device="cuda"); db = torch.empty(H, device="cuda"); dbias = torch.empty(S, H, device="cuda")
ws = torch.zeros(int(lib.bf_resln_bwd_workspace_bytes(S, M, H)), dtype=torch.uint8, device="cuda")
def fwd(k): return lib.bf_resln_fwd_keep(h.data_ptr(), r.data_ptr(), BF_BF16, g.data_ptr(), b.data_ptr(), 0, S, M, H, 1e-12, 0.1, 7, 3, 5, z.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), k, st)
def bwd(k): return lib.bf_resln_bwd_keep(gy.data_ptr(), z.data_ptr(), BF_BF16, g.data_ptr(), 0, mean.data_ptr(), rstd.data_ptr(), S, M, H, 0.1, 7, 3, 5, dz.data_ptr(), dh.data_ptr(), dg.data_ptr(), db.data_ptr(), dbias.data_ptr(), ws.data_ptr(), k, st)
def run(fn, reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): assert fn() == 0
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for _ in range(5): fwd(keep.data_ptr()); bwd(keep.data_ptr())
for rnd in range(3):
    print("fwd no-keep %.4f  keep %.4f | bwd regen %.4f  keep %.4f  (ms)" % (run(lambda: fwd(None)), run(lambda: fwd(keep.data_ptr())), run(lambda: bwd(None)), run(lambda: bwd(keep.data_ptr()))))
print("HBM floor: %.4f ms (4 passes of %d MB at 6.55 TB/s)" % (4 * rows * H * 2 / 6.5517e9, rows * H * 2 // 2**20))
