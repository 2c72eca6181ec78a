"""Diagnostics for the tcgen05 contractions (not a pytest test): prints the
error structure per (row-quarter, 32-column chunk) so descriptor / layout bugs
can be told apart in ONE GPU session."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bayeformers_b200 import _lib
from bayeformers_b200._lib import BF_BF16, BF_F32

lib = _lib.load()
DEV = "cuda:0"
st = torch.cuda.current_stream().cuda_stream


def report(name, got, want):
    err = (got.double() - want.double())
    rel = float(err.norm() / want.double().norm())
    print(f"[{name}] shape={tuple(got.shape)} rel_err={rel:.3e} max_abs={float(err.abs().max()):.3e} "
          f"nan={int(torch.isnan(got).sum())} zeros={int((got == 0).sum())}")
    if rel > 1e-3:
        g, w = got[0], want[0]
        R, C = g.shape
        for r0 in range(0, min(R, 128), 32):
            row = []
            for c0 in range(0, min(C, 256), 32):
                e = (g[r0:r0 + 32, c0:c0 + 32].double() - w[r0:r0 + 32, c0:c0 + 32].double()).norm()
                n = w[r0:r0 + 32, c0:c0 + 32].double().norm()
                row.append(f"{float(e / n):7.1e}")
            print("   rows", r0, " ".join(row))
        print("   got[0,:8] ", g[0, :8].tolist())
        print("   want[0,:8]", w[0, :8].tolist())
    return rel


def run(S, M, N, K):
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(S, M, K, generator=gen).to(DEV).bfloat16()
    w = (torch.randn(S, N, K, generator=gen) * 0.1).to(DEV).bfloat16()
    gy = torch.randn(S, M, N, generator=gen).to(DEV).bfloat16()
    print(f"=== S={S} M={M} N={N} K={K}")
    y = torch.full((S, M, N), float("nan"), device=DEV)
    rc = lib.bf_linear_fwd(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), S, M, N, K, BF_BF16, BF_F32, st)
    torch.cuda.synchronize()
    print("fwd rc", rc, lib.bf_last_error() if rc else "")
    report("fwd  (K-major x K-major)", y, torch.einsum("smk,snk->smn", x.double(), w.double()))
    dx = torch.full((S, M, K), float("nan"), device=DEV)
    rc = lib.bf_linear_dgrad(gy.data_ptr(), w.data_ptr(), dx.data_ptr(), S, M, N, K, BF_BF16, BF_F32, st)
    torch.cuda.synchronize()
    print("dgrad rc", rc, lib.bf_last_error() if rc else "")
    report("dgrad (K-major x MN-major)", dx, torch.einsum("smn,snk->smk", gy.double(), w.double()))
    dw = torch.full((S, N, K), float("nan"), device=DEV)
    rc = lib.bf_linear_wgrad(gy.data_ptr(), x.data_ptr(), dw.data_ptr(), S, M, N, K, BF_BF16, st)
    torch.cuda.synchronize()
    print("wgrad rc", rc, lib.bf_last_error() if rc else "")
    report("wgrad (MN-major x MN-major)", dw, torch.einsum("smn,smk->snk", gy.double(), x.double()))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "sm100:", lib.bf_device_is_sm100())
    run(1, 128, 256, 64)
    run(1, 128, 256, 256)
    run(2, 256, 512, 128)
    run(1, 200, 136, 72)
    # timing of the BERT-base shapes (S=4, B=32 -> M=4096) against cuBLAS bmm on the same bf16 operands
    def t_ms(fn, it=10):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it

    for (S, M, N, K) in [(4, 4096, 768, 768), (4, 4096, 3072, 768), (4, 4096, 768, 3072), (4, 16384, 3072, 768), (4, 1024, 768, 768)]:
        x = torch.randn(S, M, K, device=DEV).bfloat16(); w = torch.randn(S, N, K, device=DEV).bfloat16()
        gy = torch.randn(S, M, N, device=DEV).bfloat16(); b = torch.randn(S, N, device=DEV)
        y = torch.empty(S, M, N, device=DEV, dtype=torch.bfloat16); dx = torch.empty(S, M, K, device=DEV, dtype=torch.bfloat16)
        rho = torch.full((N, K), -5.0, device=DEV); g_rho = torch.empty(N, K, device=DEV)
        ws = torch.empty(lib.bf_linear_wgrad_fused_workspace_bytes(S, M, N, K, 0), dtype=torch.uint8, device=DEV)
        fl = 2 * S * M * N * K / 1e9
        r = {}
        r["fwd"] = t_ms(lambda: lib.bf_linear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), S, M, N, K, BF_BF16, BF_BF16, st))
        r["dgrad"] = t_ms(lambda: lib.bf_linear_dgrad(gy.data_ptr(), w.data_ptr(), dx.data_ptr(), S, M, N, K, BF_BF16, BF_BF16, st))
        r["wgrad_fused"] = t_ms(lambda: lib.bf_linear_wgrad_fused(gy.data_ptr(), x.data_ptr(), S, M, N, K, BF_BF16, None, rho.data_ptr(), 2, None, None,
                                                                  0.5, 1.0, 1.0, None, None, 7, 0, 1, None, None, g_rho.data_ptr(), 0, ws.data_ptr(), st))
        r["cublas_fwd"] = t_ms(lambda: torch.baddbmm(b[:, None, :].bfloat16(), x, w.transpose(1, 2)))
        r["cublas_dgrad"] = t_ms(lambda: torch.bmm(gy, w))
        r["cublas_wgrad"] = t_ms(lambda: torch.bmm(gy.transpose(1, 2), x))
        print(f"S={S} M={M} N={N} K={K}: " + "  ".join(f"{k} {v*1e3:.0f}us {fl/v:.0f}TF" for k, v in r.items()), flush=True)
