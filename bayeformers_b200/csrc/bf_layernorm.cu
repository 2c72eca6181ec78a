// S-sample LayerNorm forward / backward (sm_100a, HBM-bound).
//
// Row A10 of SURVEY.md section 8: the Bayesian LayerNorm the north_star names does
// not exist in the reference snapshot; it is specified by analogy with bnn.Linear as
//     gamma_s, beta_s ~ Gaussian(mu, rho)  (bayeformers/nn/parameters/gaussian.py:90-101)
//     y[s] = F.layer_norm(x[s], (H,), gamma_s, beta_s, eps)
// These kernels are that F.layer_norm and its autograd for all S samples at once
// (per-sample affine, rows [s*M, (s+1)*M) use sample s).  With S == 1 and a shared
// affine they also serve the frequentist LayerNorms of the host model.
//
// One warp owns one row: the row lives in registers as packed 16-byte chunks
// (H = 256*C elements, lane l holds elements [c*256 + l*8, +8) of every chunk), so
// x / gy are read exactly once and y / dx written once.  Statistics are two-pass
// in registers (mean, then centred sum of squares) with warp-shuffle trees.
// The affine gradients accumulate in registers over all rows a warp visits, then
// block tree -> per-block partial -> fixed-order final pass by the last block of
// each sample (no float atomics: run-to-run deterministic).
#include "bf_common.cuh"

namespace {

constexpr int kLnThreads = 384;  // backward: 1 block of 12 warps per SM (<= 168 registers per thread), 2 rows in flight per warp
constexpr int kLnWarps = kLnThreads / 32;
constexpr int kLnFwdThreads = 256;  // forward: 4 blocks of 8 warps per SM (<= 64 registers per thread)
constexpr int kLnFwdWarps = kLnFwdThreads / 32;
constexpr int kLnFwdBlocksPerSm = 4;

template <typename T>
struct Pack8;  // 8 consecutive elements <-> fp32
template <>
struct Pack8<__nv_bfloat16> {
    uint4 u;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = __ldcs(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void get(float (&v)[8]) const {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint4 o;
        uint32_t* w = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&b);
        }
        __stcs(reinterpret_cast<uint4*>(p), o);
    }
};
template <>
struct Pack8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) {
        a = __ldcs(reinterpret_cast<const float4*>(p));
        b = __ldcs(reinterpret_cast<const float4*>(p) + 1);
    }
    __device__ __forceinline__ void get(float (&v)[8]) const {
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
        __stcs(reinterpret_cast<float4*>(p) + 1, make_float4(v[4], v[5], v[6], v[7]));
    }
};

__device__ __forceinline__ void ld8f(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}

// ------------------------------------------------------------------ forward
template <typename T, int C>
__global__ void __launch_bounds__(kLnFwdThreads, kLnFwdBlocksPerSm) layernorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, T* __restrict__ y,
                                                                   float* __restrict__ mean_out,
                                                                   float* __restrict__ rstd_out, int64_t M,
                                                                   int64_t affine_stride, float eps) {
    constexpr int H = 256 * C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.y;
    const float* g = gamma + (int64_t)s * affine_stride;
    const float* b = beta ? beta + (int64_t)s * affine_stride : nullptr;
    const int64_t row0 = (int64_t)s * M;
    for (int64_t m = (int64_t)blockIdx.x * kLnFwdWarps + warp; m < M; m += (int64_t)gridDim.x * kLnFwdWarps) {
        const T* xr = x + (row0 + m) * H;
        Pack8<T> px[C];
#pragma unroll
        for (int c = 0; c < C; ++c) px[c].load(xr + c * 256 + lane * 8);
        float sum = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float v[8];
            px[c].get(v);
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += v[j];
        }
        const float mean = bf_warp_sum(sum) * (1.0f / H);
        float sq = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float v[8];
            px[c].get(v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = v[j] - mean;
                sq = fmaf(d, d, sq);
            }
        }
        const float rstd = rsqrtf(bf_warp_sum(sq) * (1.0f / H) + eps);
        if (lane == 0) {
            mean_out[row0 + m] = mean;
            rstd_out[row0 + m] = rstd;
        }
        T* yr = y + (row0 + m) * H;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float v[8], gv[8], o[8];
            px[c].get(v);
            ld8f(g + c * 256 + lane * 8, gv);
            if (b) {
                float bv[8];
                ld8f(b + c * 256 + lane * 8, bv);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaf((v[j] - mean) * rstd, gv[j], bv[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd * gv[j];
            }
            Pack8<T>::store(yr + c * 256 + lane * 8, o);
        }
    }
}

// ------------------------------------------------------------------ backward
// workspace: [S counters, padded to 256 B][S][nblk][2][H] partials
template <typename T, int C>
__global__ void __launch_bounds__(kLnThreads) layernorm_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ x,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ mean_in,
                                                                   const float* __restrict__ rstd_in, T* __restrict__ dx,
                                                                   float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                   float* __restrict__ partial,
                                                                   unsigned int* __restrict__ counters, int64_t M,
                                                                   int64_t affine_stride) {
    constexpr int H = 256 * C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.y, nblk = gridDim.x;
    const float* g = gamma + (int64_t)s * affine_stride;
    const int64_t row0 = (int64_t)s * M;
    float acc_g[C][8], acc_b[C][8];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc_g[c][j] = acc_b[c][j] = 0.0f;

    const int64_t m_step = (int64_t)nblk * kLnWarps;
    int64_t m = (int64_t)blockIdx.x * kLnWarps + warp;
    Pack8<T> nx[C], ng[C];  // software pipeline: the next row's loads are in flight while this row is reduced
    if (m < M) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            nx[c].load(x + (row0 + m) * H + c * 256 + lane * 8);
            ng[c].load(gy + (row0 + m) * H + c * 256 + lane * 8);
        }
    }
    float mean_n = 0.0f, rstd_n = 0.0f;  // the row statistics travel one iteration ahead as well (ncu: the dependent
    if (m < M) mean_n = __ldg(mean_in + row0 + m), rstd_n = __ldg(rstd_in + row0 + m);  // load cost ~20 % of the samples)
    for (; m < M; m += m_step) {
        Pack8<T> px[C], pg[C];
#pragma unroll
        for (int c = 0; c < C; ++c) px[c] = nx[c], pg[c] = ng[c];
        const float mean = mean_n, rstd = rstd_n;
        if (m + m_step < M) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                nx[c].load(x + (row0 + m + m_step) * H + c * 256 + lane * 8);
                ng[c].load(gy + (row0 + m + m_step) * H + c * 256 + lane * 8);
            }
            mean_n = __ldg(mean_in + row0 + m + m_step), rstd_n = __ldg(rstd_in + row0 + m + m_step);
        }
        float s1 = 0.0f, s2 = 0.0f;  // sum a, sum a*xhat with a = gy*gamma
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float xv[8], gv[8], gm[8];
            px[c].get(xv);
            pg[c].get(gv);
            ld8f(g + c * 256 + lane * 8, gm);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xh = (xv[j] - mean) * rstd;
                const float a = gv[j] * gm[j];
                s1 += a;
                s2 = fmaf(a, xh, s2);
                acc_g[c][j] = fmaf(gv[j], xh, acc_g[c][j]);
                acc_b[c][j] += gv[j];
            }
        }
        const float c1 = bf_warp_sum(s1) * (1.0f / H), c2 = bf_warp_sum(s2) * (1.0f / H);
        T* dr = dx + (row0 + m) * H;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float xv[8], gv[8], gm[8], o[8];
            px[c].get(xv);
            pg[c].get(gv);
            ld8f(g + c * 256 + lane * 8, gm);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xh = (xv[j] - mean) * rstd;
                o[j] = rstd * (gv[j] * gm[j] - c1 - xh * c2);
            }
            Pack8<T>::store(dr + c * 256 + lane * 8, o);
        }
    }

    // ---- block reduction: warps 4.. fold into warps 0..3 four at a time (32 KiB of shared memory at H = 1024),
    // then warps 1..3 fold into warp 0.  Fixed order -> deterministic.
    constexpr int kRedRows = 4;
    __shared__ float red[kRedRows][2 * H];
    __shared__ bool is_last;
#pragma unroll 1
    for (int base = kRedRows; base < kLnWarps + kRedRows; base += kRedRows) {
        // round `base`: writers are warps [base, base+4) -- or, in the last round, warps [1, 4) folding into warp 0
        const bool last = base >= kLnWarps;
        const int w_lo = last ? 1 : base, w_hi = last ? kRedRows : base + kRedRows;
        if (warp >= w_lo && warp < w_hi) {
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    red[warp - w_lo][c * 256 + lane * 8 + j] = acc_g[c][j];
                    red[warp - w_lo][H + c * 256 + lane * 8 + j] = acc_b[c][j];
                }
        }
        __syncthreads();
        if (!last) {
            if (warp < kRedRows) {
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        acc_g[c][j] += red[warp][c * 256 + lane * 8 + j];
                        acc_b[c][j] += red[warp][H + c * 256 + lane * 8 + j];
                    }
            }
        } else if (warp == 0) {
#pragma unroll 1
            for (int r = 0; r < kRedRows - 1; ++r)
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        acc_g[c][j] += red[r][c * 256 + lane * 8 + j];
                        acc_b[c][j] += red[r][H + c * 256 + lane * 8 + j];
                    }
        }
        __syncthreads();
    }
    float* my_part = partial + ((int64_t)s * nblk + blockIdx.x) * 2 * H;
    if (warp == 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float* pg = my_part + c * 256 + lane * 8;
            float* pb = my_part + H + c * 256 + lane * 8;
            *reinterpret_cast<float4*>(pg) = make_float4(acc_g[c][0], acc_g[c][1], acc_g[c][2], acc_g[c][3]);
            *reinterpret_cast<float4*>(pg + 4) = make_float4(acc_g[c][4], acc_g[c][5], acc_g[c][6], acc_g[c][7]);
            *reinterpret_cast<float4*>(pb) = make_float4(acc_b[c][0], acc_b[c][1], acc_b[c][2], acc_b[c][3]);
            *reinterpret_cast<float4*>(pb + 4) = make_float4(acc_b[c][4], acc_b[c][5], acc_b[c][6], acc_b[c][7]);
        }
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(counters + s, 1u);
        is_last = (done == (unsigned int)nblk - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- final fixed-order pass by the last block of this sample ----
    const volatile float* base = partial + (int64_t)s * nblk * 2 * H;
    for (int col = threadIdx.x; col < 2 * H; col += kLnThreads) {
        float t = 0.0f;
        for (int k = 0; k < nblk; ++k) t += base[(int64_t)k * 2 * H + col];
        if (col < H) dgamma[(int64_t)s * H + col] = t;
        else if (dbeta) dbeta[(int64_t)s * H + col - H] = t;
    }
    if (threadIdx.x == 0) counters[s] = 0u;
}

inline int ln_blocks(int64_t S, int64_t M) {
    const int64_t sms = bf_num_sms();
    int64_t per = sms / S;  // 1 block of 16 warps per SM (register budget), spread over the samples
    if (per < 1) per = 1;
    const int64_t need = (M + kLnWarps - 1) / kLnWarps;
    if (per > need) per = need;
    return (int)(per < 1 ? 1 : per);
}

template <typename T, int C>
int launch_fwd_c(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int64_t S,
                 int64_t M, int64_t astride, float eps, cudaStream_t st) {
    int64_t need = (M + kLnFwdWarps - 1) / kLnFwdWarps;
    int64_t cap = (int64_t)bf_num_sms() * kLnFwdBlocksPerSm / S;  // one full wave of resident blocks over all samples
    if (cap < 1) cap = 1;
    if (need > cap) need = cap;
    dim3 grid((unsigned)(need < 1 ? 1 : need), (unsigned)S);
    layernorm_fwd_kernel<T, C><<<grid, kLnFwdThreads, 0, st>>>(reinterpret_cast<const T*>(x), gamma, beta,
                                                             reinterpret_cast<T*>(y), mean, rstd, M, astride, eps);
    return 0;
}

template <typename T, int C>
int launch_bwd_c(const void* gy, const void* x, const float* gamma, const float* mean, const float* rstd, void* dx,
                 float* dgamma, float* dbeta, void* ws, int64_t S, int64_t M, int64_t astride, cudaStream_t st) {
    const int nblk = ln_blocks(S, M);
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
    float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + ((S * 4 + 255) / 256) * 256);
    dim3 grid((unsigned)nblk, (unsigned)S);
    layernorm_bwd_kernel<T, C><<<grid, kLnThreads, 0, st>>>(reinterpret_cast<const T*>(gy), reinterpret_cast<const T*>(x),
                                                             gamma, mean, rstd, reinterpret_cast<T*>(dx), dgamma, dbeta,
                                                             partial, counters, M, astride);
    return 0;
}

#define BF_LN_DISPATCH(FN, T, ...)                    \
    switch (H / 256) {                                \
        case 1: FN<T, 1>(__VA_ARGS__); break;         \
        case 2: FN<T, 2>(__VA_ARGS__); break;         \
        case 3: FN<T, 3>(__VA_ARGS__); break;         \
        case 4: FN<T, 4>(__VA_ARGS__); break;         \
        default: bf_set_error("LayerNorm: H/256 must be 1..4"); return BF_ERR_UNSUPPORTED; \
    }

}  // namespace

extern "C" int bf_layernorm_supported(int64_t H) { return (H % 256 == 0 && H >= 256 && H <= 1024) ? 1 : 0; }

extern "C" int bf_layernorm_fwd(const void* x, int32_t dtype, const float* gamma, const float* beta, int64_t affine_stride,
                                int64_t S, int64_t M, int64_t H, float eps, void* y, float* mean, float* rstd,
                                void* stream) {
    BF_CHECK_ARG(x && gamma && y && mean && rstd, "null pointer");
    BF_CHECK_ARG(dtype == BF_F32 || dtype == BF_BF16, "bad dtype");
    BF_CHECK_ARG(S >= 1 && M >= 0 && S <= 65535, "bad S or M");
    BF_CHECK_ARG(bf_layernorm_supported(H), "H must be 256, 512, 768 or 1024");
    BF_CHECK_ARG(affine_stride == 0 || affine_stride >= H, "bad affine stride");
    if (M == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == BF_BF16) {
        BF_LN_DISPATCH(launch_fwd_c, __nv_bfloat16, x, gamma, beta, y, mean, rstd, S, M, affine_stride, eps, st);
    } else {
        BF_LN_DISPATCH(launch_fwd_c, float, x, gamma, beta, y, mean, rstd, S, M, affine_stride, eps, st);
    }
    BF_LAUNCH_OK();
    return 0;
}

extern "C" int64_t bf_layernorm_bwd_workspace_bytes(int64_t S, int64_t M, int64_t H) {
    if (S < 1) S = 1;
    return ((S * 4 + 255) / 256) * 256 + S * (int64_t)ln_blocks(S, M < 1 ? 1 : M) * 2 * H * (int64_t)sizeof(float);
}

extern "C" int bf_layernorm_bwd(const void* gy, const void* x, int32_t dtype, const float* gamma, int64_t affine_stride,
                                const float* mean, const float* rstd, int64_t S, int64_t M, int64_t H, void* dx,
                                float* dgamma, float* dbeta, void* workspace, void* stream) {
    BF_CHECK_ARG(gy && x && gamma && mean && rstd && dx && dgamma && workspace, "null pointer");
    BF_CHECK_ARG(dtype == BF_F32 || dtype == BF_BF16, "bad dtype");
    BF_CHECK_ARG(S >= 1 && M >= 1 && S <= 65535, "bad S or M");
    BF_CHECK_ARG(bf_layernorm_supported(H), "H must be 256, 512, 768 or 1024");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == BF_BF16) {
        BF_LN_DISPATCH(launch_bwd_c, __nv_bfloat16, gy, x, gamma, mean, rstd, dx, dgamma, dbeta, workspace, S, M,
                       affine_stride, st);
    } else {
        BF_LN_DISPATCH(launch_bwd_c, float, gy, x, gamma, mean, rstd, dx, dgamma, dbeta, workspace, S, M, affine_stride,
                       st);
    }
    BF_LAUNCH_OK();
    return 0;
}
