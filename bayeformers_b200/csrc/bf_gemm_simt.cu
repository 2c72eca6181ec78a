// fp32 FFMA contractions for the reference-precision ("parity") mode, plus the
// bias-gradient column sum used by both modes.
//
// Replaces F.linear(input, weight, bias) (bayeformers/nn/layers/linear.py:104)
// and the mm/addmm calls its autograd makes, batched over S weight samples.
// These kernels exist so that injected-eps logits can be compared with the
// reference at 1e-5 (TF32/bf16 tensor-core products cannot meet that); the
// throughput path is bf_gemm_tc.cu.
#include "bf_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
constexpr int kGemmThreads = (BM / TM) * (BN / TN);  // 256

// C[b][i][j] = sum_k A(b,i,k) * B(b,k,j) (+ bias[b][j]);  A(b,i,k) = a[b*a_bs + i*a_rs + k*a_cs] etc.
struct GemmParams {
    const float* a;
    const float* b;
    const float* bias;
    float* c;
    int64_t I, J, Kd;
    int64_t a_bs, a_rs, a_cs;
    int64_t b_bs, b_rs, b_cs;
    int64_t c_bs, bias_bs;
};

__global__ void __launch_bounds__(kGemmThreads) gemm_f32_kernel(const GemmParams p) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int bz = blockIdx.z;
    const float* __restrict__ A = p.a + (int64_t)bz * p.a_bs;
    const float* __restrict__ B = p.b + (int64_t)bz * p.b_bs;
    float* __restrict__ C = p.c + (int64_t)bz * p.c_bs;
    const int64_t i_base = (int64_t)blockIdx.y * BM, j_base = (int64_t)blockIdx.x * BN;
    const int t = threadIdx.x;
    const int ti = t / (BN / TN), tj = t % (BN / TN);

    // loader mappings: walk the contiguous axis with consecutive threads
    const bool a_k_contig = (p.a_cs == 1);
    const bool b_j_contig = (p.b_cs == 1);

    float acc[TM][TN];
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[r][c] = 0.0f;

    for (int64_t k0 = 0; k0 < p.Kd; k0 += BK) {
        // ---- A tile: BM x BK = 1024 elements, 4 per thread
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = e * kGemmThreads + t;
            int i, k;
            if (a_k_contig) {
                i = idx / BK, k = idx % BK;
            } else {
                k = idx / BM, i = idx % BM;
            }
            const int64_t gi = i_base + i, gk = k0 + k;
            As[k][i] = (gi < p.I && gk < p.Kd) ? __ldg(A + gi * p.a_rs + gk * p.a_cs) : 0.0f;
        }
        // ---- B tile: BK x BN
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = e * kGemmThreads + t;
            int k, j;
            if (b_j_contig) {
                k = idx / BN, j = idx % BN;
            } else {
                j = idx / BK, k = idx % BK;
            }
            const int64_t gk = k0 + k, gj = j_base + j;
            Bs[k][j] = (gk < p.Kd && gj < p.J) ? __ldg(B + gk * p.b_rs + gj * p.b_cs) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[TM], bv[TN];
#pragma unroll
            for (int r = 0; r < TM; ++r) av[r] = As[k][ti * TM + r];
#pragma unroll
            for (int c = 0; c < TN; ++c) bv[c] = Bs[k][tj * TN + c];
#pragma unroll
            for (int r = 0; r < TM; ++r)
#pragma unroll
                for (int c = 0; c < TN; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
        }
        __syncthreads();
    }
    const float* bias = p.bias ? p.bias + (int64_t)bz * p.bias_bs : nullptr;
#pragma unroll
    for (int r = 0; r < TM; ++r) {
        const int64_t gi = i_base + ti * TM + r;
        if (gi >= p.I) continue;
#pragma unroll
        for (int c = 0; c < TN; ++c) {
            const int64_t gj = j_base + tj * TN + c;
            if (gj >= p.J) continue;
            float v = acc[r][c];
            if (bias) v += __ldg(bias + gj);
            C[gi * p.J + gj] = v;
        }
    }
}

int launch_gemm(const GemmParams& p, int64_t S, cudaStream_t st) {
    if (p.I == 0 || p.J == 0 || S == 0) return 0;
    dim3 grid((unsigned)((p.J + BN - 1) / BN), (unsigned)((p.I + BM - 1) / BM), (unsigned)S);
    gemm_f32_kernel<<<grid, kGemmThreads, 0, st>>>(p);
    return 0;
}

// ---- bias gradient: db[s][j] = sum_m gy[s][m][j] ---------------------------
// Two deterministic stages in one launch: every block reduces a slab of rows for 32*V columns
// into a partial; the last block to finish a (sample, column-block) adds the partials in
// slab order (self-resetting counter, no float atomics).  V = elements per lane and load:
// 8 bf16 / 4 fp32 (16-byte loads, N % V == 0, 16 B aligned rows) or 2 (scalar, any shape).
constexpr int kBiasThreads = 256;
constexpr int kBiasWarps = kBiasThreads / 32;
constexpr int kBiasRowsInFlight = 4;

template <typename T, int V>
struct BiasLoad;
template <>
struct BiasLoad<__nv_bfloat16, 8> {
    static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 u = __ldcs(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
};
template <>
struct BiasLoad<float, 4> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
        const float4 u = __ldcs(reinterpret_cast<const float4*>(p));
        v[0] = u.x, v[1] = u.y, v[2] = u.z, v[3] = u.w;
    }
};

template <typename T, int V, bool VEC>
__global__ void __launch_bounds__(kBiasThreads) bias_grad_kernel(const T* __restrict__ gy, float* __restrict__ db,
                                                                 float* __restrict__ partial,
                                                                 unsigned int* __restrict__ counters, int64_t M,
                                                                 int64_t N, int64_t rows_per_slab) {
    constexpr int kCols = 32 * V;
    const int s = blockIdx.y, slab = blockIdx.z, n_slabs = gridDim.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * kCols + lane * V;
    const T* base = gy + (int64_t)s * M * N;
    const int64_t m_lo = (int64_t)slab * rows_per_slab;
    const int64_t m_hi = m_lo + rows_per_slab < M ? m_lo + rows_per_slab : M;
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.0f;
    int64_t m = m_lo + warp;
    if constexpr (VEC) {
        if (c0 < N) {  // N % V == 0: a lane's V columns are all inside or all outside
            for (; m + (kBiasRowsInFlight - 1) * kBiasWarps < m_hi; m += kBiasRowsInFlight * kBiasWarps) {
                float v[kBiasRowsInFlight][V];
#pragma unroll
                for (int u = 0; u < kBiasRowsInFlight; ++u) BiasLoad<T, V>::ld(base + (m + u * kBiasWarps) * N + c0, v[u]);
#pragma unroll
                for (int u = 0; u < kBiasRowsInFlight; ++u)
#pragma unroll
                    for (int j = 0; j < V; ++j) acc[j] += v[u][j];
            }
            for (; m < m_hi; m += kBiasWarps) {
                float v[V];
                BiasLoad<T, V>::ld(base + m * N + c0, v);
#pragma unroll
                for (int j = 0; j < V; ++j) acc[j] += v[j];
            }
        }
    } else {
        for (; m < m_hi; m += kBiasWarps) {
            const T* row = base + m * N + c0;
#pragma unroll
            for (int j = 0; j < V; ++j)
                if (c0 + j < N) acc[j] += bf_ld_as_float(row + j);
        }
    }
    __shared__ float red[kBiasWarps][kCols];
    __shared__ bool is_last;
#pragma unroll
    for (int j = 0; j < V; ++j) red[warp][lane * V + j] = acc[j];
    __syncthreads();
    const int64_t cb = blockIdx.x;
    const int64_t n_cb = gridDim.x;
    for (int c = threadIdx.x; c < kCols; c += kBiasThreads) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kBiasWarps; ++w) t += red[w][c];
        partial[(((int64_t)s * n_cb + cb) * n_slabs + slab) * kCols + c] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(counters + s * n_cb + cb, 1u);
        is_last = (done == (unsigned int)n_slabs - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int cc = threadIdx.x; cc < kCols; cc += kBiasThreads) {
        const volatile float* pp = partial + ((int64_t)s * n_cb + cb) * n_slabs * kCols + cc;
        float t = 0.0f;
        for (int k = 0; k < n_slabs; ++k) t += pp[(int64_t)k * kCols];
        const int64_t c = cb * kCols + cc;
        if (c < N) db[(int64_t)s * N + c] = t;
    }
    if (threadIdx.x == 0) counters[s * n_cb + cb] = 0u;
}

inline void bias_grid(int64_t S, int64_t M, int64_t N, int cols, int& n_cb, int& n_slabs, int64_t& rows_per_slab) {
    n_cb = (int)((N + cols - 1) / cols);
    const int64_t target = (int64_t)bf_num_sms() * 4;
    int64_t slabs = target / (S * n_cb);
    const int64_t max_slabs = (M + 63) / 64;
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    if (slabs > 65535) slabs = 65535;
    rows_per_slab = (M + slabs - 1) / slabs;
    n_slabs = (int)((M + rows_per_slab - 1) / rows_per_slab);
}

inline int64_t bias_ws_bytes(int64_t S, int64_t M, int64_t N, int cols) {
    int n_cb, n_slabs;
    int64_t rps;
    bias_grid(S, M, N, cols, n_cb, n_slabs, rps);
    // [counters: S*n_cb uint32, padded to 256 B][partials: S*n_cb*n_slabs*cols floats]
    const int64_t cnt = ((S * n_cb * 4 + 255) / 256) * 256;
    return cnt + S * n_cb * (int64_t)n_slabs * cols * 4;
}

template <typename T, int V, bool VEC>
void launch_bias(const void* gy, float* db, int64_t S, int64_t M, int64_t N, void* workspace, cudaStream_t st) {
    int n_cb, n_slabs;
    int64_t rps;
    bias_grid(S, M < 1 ? 1 : M, N, 32 * V, n_cb, n_slabs, rps);
    const int64_t cnt = ((S * n_cb * 4 + 255) / 256) * 256;
    unsigned int* counters = reinterpret_cast<unsigned int*>(workspace);
    float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + cnt);
    dim3 grid((unsigned)n_cb, (unsigned)S, (unsigned)n_slabs);
    bias_grad_kernel<T, V, VEC><<<grid, kBiasThreads, 0, st>>>(reinterpret_cast<const T*>(gy), db, partial, counters, M,
                                                                N, rps);
}

}  // namespace

int bf_linear_fwd_f32(const float* x, const float* w, const float* bias, float* y, int64_t S, int64_t M, int64_t N,
                      int64_t K, cudaStream_t st) {
    GemmParams p{};
    p.a = x, p.b = w, p.bias = bias, p.c = y;
    p.I = M, p.J = N, p.Kd = K;
    p.a_bs = M * K, p.a_rs = K, p.a_cs = 1;
    p.b_bs = N * K, p.b_rs = 1, p.b_cs = K;  // B(k,j) = w[j][k]
    p.c_bs = M * N, p.bias_bs = N;
    return launch_gemm(p, S, st);
}

int bf_linear_dgrad_f32(const float* gy, const float* w, float* dx, int64_t S, int64_t M, int64_t N, int64_t K,
                        cudaStream_t st) {
    GemmParams p{};
    p.a = gy, p.b = w, p.bias = nullptr, p.c = dx;
    p.I = M, p.J = K, p.Kd = N;
    p.a_bs = M * N, p.a_rs = N, p.a_cs = 1;
    p.b_bs = N * K, p.b_rs = K, p.b_cs = 1;  // B(n,k) = w[n][k]
    p.c_bs = M * K, p.bias_bs = 0;
    return launch_gemm(p, S, st);
}

int bf_linear_wgrad_f32(const float* gy, const float* x, float* dw, int64_t S, int64_t M, int64_t N, int64_t K,
                        cudaStream_t st) {
    GemmParams p{};
    p.a = gy, p.b = x, p.bias = nullptr, p.c = dw;
    p.I = N, p.J = K, p.Kd = M;
    p.a_bs = M * N, p.a_rs = 1, p.a_cs = N;  // A(n,m) = gy[m][n]
    p.b_bs = M * K, p.b_rs = K, p.b_cs = 1;  // B(m,k) = x[m][k]
    p.c_bs = N * K, p.bias_bs = 0;
    return launch_gemm(p, S, st);
}

extern "C" int64_t bf_bias_grad_workspace_bytes(int64_t S, int64_t M, int64_t N) {
    S = S < 1 ? 1 : S, M = M < 1 ? 1 : M, N = N < 1 ? 1 : N;
    int64_t b = bias_ws_bytes(S, M, N, 64);  // large enough for whichever variant bf_bias_grad picks
    const int64_t b4 = bias_ws_bytes(S, M, N, 128), b8 = bias_ws_bytes(S, M, N, 256);
    if (b4 > b) b = b4;
    if (b8 > b) b = b8;
    return b;
}

extern "C" int bf_bias_grad(const void* gy, int32_t gy_dtype, float* db, int64_t S, int64_t M, int64_t N,
                            void* workspace, void* stream) {
    BF_CHECK_ARG(gy && db && workspace, "null pointer");
    BF_CHECK_ARG(gy_dtype == BF_F32 || gy_dtype == BF_BF16, "bad dtype");
    if (S == 0 || N == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool al16 = (reinterpret_cast<uintptr_t>(gy) & 15u) == 0;
    if (gy_dtype == BF_BF16) {
        if (al16 && N % 8 == 0) launch_bias<__nv_bfloat16, 8, true>(gy, db, S, M, N, workspace, st);
        else launch_bias<__nv_bfloat16, 2, false>(gy, db, S, M, N, workspace, st);
    } else {
        if (al16 && N % 4 == 0) launch_bias<float, 4, true>(gy, db, S, M, N, workspace, st);
        else launch_bias<float, 2, false>(gy, db, S, M, N, workspace, st);
    }
    BF_LAUNCH_OK();
    return 0;
}
