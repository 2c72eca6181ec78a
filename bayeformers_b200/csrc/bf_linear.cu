// C-ABI entry points of the S-sample Linear contractions: argument checks and
// dispatch between the fp32 FFMA kernels (bf_gemm_simt.cu) and the tcgen05
// kernels (bf_gemm_tc.cu).  Contract: include/bayeformers_b200.h.
#include "bf_common.cuh"

int bf_linear_fwd_f32(const float*, const float*, const float*, float*, int64_t, int64_t, int64_t, int64_t, cudaStream_t);
int bf_linear_dgrad_f32(const float*, const float*, float*, int64_t, int64_t, int64_t, int64_t, cudaStream_t);
int bf_linear_wgrad_f32(const float*, const float*, float*, int64_t, int64_t, int64_t, int64_t, cudaStream_t);
int bf_linear_fwd_bf16(const void*, const void*, const float*, void*, int64_t, int64_t, int64_t, int64_t, int32_t, cudaStream_t);
int bf_linear_dgrad_bf16(const void*, const void*, void*, int64_t, int64_t, int64_t, int64_t, int32_t, int, cudaStream_t);
int bf_linear_wgrad_bf16(const void*, const void*, float*, int64_t, int64_t, int64_t, int64_t, cudaStream_t);
int64_t bf_wgrad_fused_workspace_bytes_impl(int64_t S, int64_t M, int64_t N, int64_t K, int with_mu);
int bf_linear_wgrad_fused_bf16(const void*, const void*, int64_t, int64_t, int64_t, int64_t, const float*, const float*,
                               int32_t, const float*, const float*, float, float, float, const float*, const float*,
                               uint64_t, uint32_t, uint32_t, const float*, float*, float*, int32_t, void*, cudaStream_t);

#define BF_CHECK_SHAPE()                                                        \
    BF_CHECK_ARG(S >= 1 && M >= 1 && N >= 1 && K >= 1, "S, M, N, K must be >= 1"); \
    BF_CHECK_ARG(dtype == BF_F32 || dtype == BF_BF16, "bad dtype")

extern "C" int bf_linear_fwd(const void* x, const void* w, const float* bias, void* y, int64_t S, int64_t M, int64_t N,
                             int64_t K, int32_t dtype, int32_t y_dtype, void* stream) {
    BF_CHECK_ARG(x && w && y, "null pointer");
    BF_CHECK_SHAPE();
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == BF_F32) {
        BF_CHECK_ARG(y_dtype == BF_F32, "fp32 mode writes fp32");
        int rc = bf_linear_fwd_f32((const float*)x, (const float*)w, bias, (float*)y, S, M, N, K, st);
        if (rc) return rc;
        BF_LAUNCH_OK();
        return 0;
    }
    BF_CHECK_ARG(y_dtype == BF_F32 || y_dtype == BF_BF16, "bad y_dtype");
    return bf_linear_fwd_bf16(x, w, bias, y, S, M, N, K, y_dtype, st);
}

extern "C" int bf_linear_dgrad(const void* gy, const void* w, void* dx, int64_t S, int64_t M, int64_t N, int64_t K,
                               int32_t dtype, int32_t dx_dtype, void* stream) {
    BF_CHECK_ARG(gy && w && dx, "null pointer");
    BF_CHECK_SHAPE();
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == BF_F32) {
        BF_CHECK_ARG(dx_dtype == BF_F32, "fp32 mode writes fp32");
        int rc = bf_linear_dgrad_f32((const float*)gy, (const float*)w, (float*)dx, S, M, N, K, st);
        if (rc) return rc;
        BF_LAUNCH_OK();
        return 0;
    }
    BF_CHECK_ARG(dx_dtype == BF_F32 || dx_dtype == BF_BF16, "bad dx_dtype");
    return bf_linear_dgrad_bf16(gy, w, dx, S, M, N, K, dx_dtype, 0, st);
}

extern "C" int bf_linear_dgrad_accumulate(const void* gy, const void* w, void* dx, int64_t S, int64_t M, int64_t N,
                                          int64_t K, int32_t dtype, int32_t dx_dtype, void* stream) {
    BF_CHECK_ARG(gy && w && dx, "null pointer");
    BF_CHECK_SHAPE();
    BF_CHECK_ARG(dtype == BF_BF16, "accumulating dgrad exists for the bf16 tensor-core path only");
    BF_CHECK_ARG(dx_dtype == BF_F32 || dx_dtype == BF_BF16, "bad dx_dtype");
    return bf_linear_dgrad_bf16(gy, w, dx, S, M, N, K, dx_dtype, 1, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int bf_linear_wgrad(const void* gy, const void* x, float* dw, int64_t S, int64_t M, int64_t N, int64_t K,
                               int32_t dtype, void* stream) {
    BF_CHECK_ARG(gy && x && dw, "null pointer");
    BF_CHECK_SHAPE();
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == BF_F32) {
        int rc = bf_linear_wgrad_f32((const float*)gy, (const float*)x, dw, S, M, N, K, st);
        if (rc) return rc;
        BF_LAUNCH_OK();
        return 0;
    }
    return bf_linear_wgrad_bf16(gy, x, dw, S, M, N, K, st);
}

extern "C" int64_t bf_linear_wgrad_fused_workspace_bytes(int64_t S, int64_t M, int64_t N, int64_t K,
                                                         int32_t with_grad_mu) {
    if (S < 1 || M < 1 || N < 1 || K < 1) return 0;
    return bf_wgrad_fused_workspace_bytes_impl(S, M, N, K, with_grad_mu);
}

extern "C" int bf_linear_wgrad_fused(const void* gy, const void* x, int64_t S, int64_t M, int64_t N, int64_t K,
                                     int32_t dtype, const float* mu, const float* rho, int32_t prior_kind,
                                     const float* prior_mu, const float* prior_rho, float pi, float sigma1,
                                     float sigma2, const float* g_logq, const float* g_logp, uint64_t seed,
                                     uint32_t step, uint32_t tensor_id, const float* eps_in, float* grad_mu,
                                     float* grad_rho, int32_t accumulate, void* workspace, void* stream) {
    BF_CHECK_ARG(gy && x && rho && grad_rho && workspace, "null pointer");
    BF_CHECK_SHAPE();
    BF_CHECK_ARG(dtype == BF_BF16, "fused wgrad exists for the bf16 tensor-core path only");
    const bool kl = g_logq || g_logp;
    BF_CHECK_ARG(!kl || mu, "KL gradient needs mu");
    BF_CHECK_ARG(!kl || prior_kind != BF_PRIOR_GAUSSIAN || prior_mu, "gaussian prior needs prior_mu");
    return bf_linear_wgrad_fused_bf16(gy, x, S, M, N, K, mu, rho, prior_kind, prior_mu, prior_rho, pi, sigma1, sigma2,
                                      g_logq, g_logp, seed, step, tensor_id, eps_in, grad_mu, grad_rho, accumulate,
                                      workspace, reinterpret_cast<cudaStream_t>(stream));
}
