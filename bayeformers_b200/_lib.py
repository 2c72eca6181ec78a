"""ctypes binding of the C-ABI library (include/bayeformers_b200.h).

There is deliberately no fallback: if the shared object is missing or a call
fails, the product raises.  The library is a plain `extern "C"` .so (no torch
types in any signature); PyTorch only supplies device pointers and the stream.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_uint32, c_uint64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libbayeformers_b200.so")

ABI_VERSION = 2
BF_F32, BF_BF16 = 0, 1
BF_PRIOR_MIXTURE, BF_PRIOR_GAUSSIAN, BF_PRIOR_NONE = 0, 1, 2
BF_OPT_GEMM_2CTA, BF_OPT_WGRAD_2CTA, BF_OPT_RESLN_BWD_STAGED, BF_OPT_SK_PREFETCH, BF_OPT_ATTN_TC, BF_OPT_GELU_POLY = 0, 1, 2, 3, 4, 5

class BfTensorDesc(ctypes.Structure):
    """`bf_tensor_desc` of include/bayeformers_b200.h (multi-tensor sample+KL)."""
    _fields_ = [("mu", c_void_p), ("rho", c_void_p), ("prior_mu", c_void_p), ("prior_rho", c_void_p),
                ("w_out", c_void_p), ("n", c_int64), ("w_stride", c_int64), ("tensor_id", c_uint32),
                ("step", c_uint32), ("prior_kind", c_int32), ("w_dtype", c_int32), ("pi", c_float),
                ("sigma1", c_float), ("sigma2", c_float), ("vec", c_int32), ("sigma", c_void_p)]


class BfOptDesc(ctypes.Structure):
    """`bf_opt_desc` of include/bayeformers_b200.h (fused clip + AdamW)."""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("n", c_int64), ("dtype", c_int32), ("vec", c_int32), ("master", c_void_p), ("sigma_out", c_void_p)]


# name -> (restype, argtypes); mirrors include/bayeformers_b200.h one to one
SIGNATURES = {
    "bf_abi_version": (c_int32, []),
    "bf_last_error": (c_char_p, []),
    "bf_device_is_sm100": (c_int32, []),
    "bf_set_step_counter": (c_int32, [c_void_p]),
    "bf_set_option": (c_int32, [c_int32, c_int32]),
    "bf_get_option": (c_int32, [c_int32]),
    "bf_philox_normal": (c_int32, [c_void_p, c_int64, c_uint64, c_uint32, c_uint32, c_uint32, c_void_p]),
    "bf_sample_kl_workspace_bytes": (c_int64, [c_int64, c_int32]),
    "bf_sample_kl_fwd": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_float, c_float, c_float,
                                   c_int64, c_int32, c_uint64, c_uint32, c_uint32, c_void_p, c_void_p, c_int32,
                                   c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "bf_softplus_fwd": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p]),
    "bf_sample_kl_multi_chunk_quads": (c_int32, []),
    "bf_sample_kl_multi_workspace_bytes": (c_int64, [c_int64]),
    "bf_sample_kl_fwd_multi": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_uint64,
                                         c_uint32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "bf_sample_kl_bwd": (c_int32, [c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                   c_float, c_float, c_float, c_void_p, c_void_p, c_int64, c_int32, c_uint64,
                                   c_uint32, c_uint32, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    "bf_linear_fwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                c_int32, c_int32, c_void_p]),
    "bf_linear_dgrad": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32,
                                  c_int32, c_void_p]),
    "bf_linear_dgrad_accumulate": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32,
                                             c_int32, c_void_p]),
    "bf_linear_wgrad": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32,
                                  c_void_p]),
    "bf_split_bf16x2": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "bf_linear_fwd_x3": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                   c_int64, c_void_p]),
    "bf_linear_dgrad_x3": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                     c_void_p]),
    "bf_linear_wgrad_x3": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                     c_void_p]),
    "bf_linear_fwd_gelu_supported": (c_int32, [c_int64, c_int64, c_int64, c_int64]),
    "bf_linear_fwd_gelu": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                     c_int64, c_void_p]),
    "bf_linear_dgrad_gelu_supported": (c_int32, [c_int64, c_int64, c_int64, c_int64]),
    "bf_linear_dgrad_gelu": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "bf_linear_dgrad_gelu_bias_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "bf_linear_dgrad_gelu_bias": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                            c_int64, c_int64, c_void_p]),
    "bf_gelu_bwd_bias_grad_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "bf_gelu_bwd_bias_grad": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                        c_void_p]),
    "bf_linear_wgrad_fused_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64, c_int64, c_int32]),
    "bf_linear_wgrad_fused": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32, c_void_p,
                                        c_void_p, c_int32, c_void_p, c_void_p, c_float, c_float, c_float, c_void_p,
                                        c_void_p, c_uint64, c_uint32, c_uint32, c_void_p, c_void_p, c_void_p,
                                        c_int32, c_void_p, c_void_p]),
    "bf_optim_chunk_elems": (c_int32, []),
    "bf_clip_adamw_workspace_bytes": (c_int64, [c_int64]),
    "bf_clip_adamw_step": (c_int32, [c_void_p, c_void_p, c_int32, c_float, c_float, c_float, c_float, c_float, c_float,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "bf_bias_grad_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "bf_layernorm_supported": (c_int32, [c_int64]),
    "bf_layernorm_fwd": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_float,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "bf_layernorm_bwd_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "bf_layernorm_bwd": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                                   c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "bf_bias_grad": (c_int32, [c_void_p, c_int32, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "bf_embedding_supported": (c_int32, [c_int64]),
    "bf_embedding_fwd": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_uint64, c_uint32,
                                   c_uint32, c_void_p, c_void_p, c_int32, c_void_p]),
    "bf_embedding_bwd_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "bf_embedding_bwd": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64,
                                   c_int64, c_uint64, c_uint32, c_uint32, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p]),
    "bf_resln_supported": (c_int32, [c_int64]),
    "bf_resln_fwd": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                               c_float, c_float, c_uint64, c_uint32, c_uint32, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    "bf_resln_bwd_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "bf_resln_bwd": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                               c_int64, c_float, c_uint64, c_uint32, c_uint32, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p]),
    "bf_resln_fwd_keep": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                    c_float, c_float, c_uint64, c_uint32, c_uint32, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "bf_resln_bwd_keep": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                                    c_int64, c_float, c_uint64, c_uint32, c_uint32, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "bf_dropout_mask": (c_int32, [c_void_p, c_int64, c_float, c_uint64, c_uint32, c_uint32, c_void_p]),
    "bf_attention_supported": (c_int32, [c_int64, c_int64]),
    "bf_attention_fwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float, c_float,
                                   c_uint64, c_uint32, c_uint32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "bf_attention_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                   c_int64, c_int64, c_float, c_float, c_uint64, c_uint32, c_uint32, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "bf_attention_bias_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "bf_attention_bwd_bias": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                        c_int64, c_int64, c_float, c_float, c_uint64, c_uint32, c_uint32, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "bf_attention_dropout_mask": (c_int32, [c_void_p, c_int64, c_int64, c_int64, c_float, c_uint64, c_uint32, c_uint32,
                                            c_void_p]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load (once) and type the library.  Raises NativeLibraryError when it has
    not been built -- there is no Python/CPU substitute for the kernels."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found: build it with `python -m bayeformers_b200.build` "
            "(the variational-layer kernels are CUDA-only; no fallback exists)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.bf_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f"ABI version mismatch: library {lib.bf_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().bf_last_error()
        raise NativeLibraryError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
