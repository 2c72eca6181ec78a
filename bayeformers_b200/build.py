"""In-tree build of libbayeformers_b200.so (sm_100a only, hand-written CUDA).

    python -m bayeformers_b200.build            # build if sources are newer
    python -m bayeformers_b200.build --force

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the
working tree to the GPU box.  No torch / pybind in the library: it is a plain
C-ABI shared object (include/bayeformers_b200.h) loaded with ctypes.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "csrc", "_obj")
LIB = os.path.join(PKG, "libbayeformers_b200.so")
SOURCES = ["bf_api.cu", "bf_sample_kl.cu", "bf_gemm_simt.cu", "bf_gemm_tc.cu", "bf_gemm_tc2.cu", "bf_gemm_act.cu", "bf_wgrad_tc.cu", "bf_linear.cu", "bf_layernorm.cu", "bf_optim.cu", "bf_resln.cu", "bf_embedding.cu", "bf_attention.cu", "bf_attention_tc.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _newer(src_files, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(f) > t for f in src_files)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, "bf_common.cuh"), os.path.join(CSRC, "bf_tc.cuh"), os.path.join(PKG, "..", "include", "bayeformers_b200.h")]
    if not force and not _newer(deps, LIB):
        return LIB
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
