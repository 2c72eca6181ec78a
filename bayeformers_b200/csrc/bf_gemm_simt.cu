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
// Two deterministic stages in one launch: every block reduces a slab of rows for 64 columns
// into a partial; the last block to finish a (sample, column-block) adds the partials in
// slab order (self-resetting counter, no float atomics).
constexpr int kColsPerBlock = 64;
constexpr int kBiasThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kBiasThreads) bias_grad_kernel(const T* __restrict__ gy, float* __restrict__ db,
                                                                 float* __restrict__ partial,
                                                                 unsigned int* __restrict__ counters, int64_t M,
                                                                 int64_t N, int64_t rows_per_slab) {
    const int s = blockIdx.y, slab = blockIdx.z, n_slabs = gridDim.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * kColsPerBlock + lane * 2;
    const T* base = gy + (int64_t)s * M * N;
    const int64_t m_lo = (int64_t)slab * rows_per_slab;
    const int64_t m_hi = m_lo + rows_per_slab < M ? m_lo + rows_per_slab : M;
    float a0 = 0.0f, a1 = 0.0f;
    const bool ok0 = c0 < N, ok1 = c0 + 1 < N;
    int64_t m = m_lo + warp;
    for (; m + 3 * (kBiasThreads / 32) < m_hi; m += 4 * (kBiasThreads / 32)) {  // 4 rows in flight per warp
        float v0[4], v1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const T* row = base + (m + u * (kBiasThreads / 32)) * N + c0;
            v0[u] = ok0 ? bf_ld_as_float(row) : 0.0f;
            v1[u] = ok1 ? bf_ld_as_float(row + 1) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) a0 += v0[u], a1 += v1[u];
    }
    for (; m < m_hi; m += kBiasThreads / 32) {
        const T* row = base + m * N + c0;
        if (ok0) a0 += bf_ld_as_float(row);
        if (ok1) a1 += bf_ld_as_float(row + 1);
    }
    __shared__ float red[kBiasThreads / 32][kColsPerBlock];
    __shared__ bool is_last;
    red[warp][lane * 2] = a0;
    red[warp][lane * 2 + 1] = a1;
    __syncthreads();
    const int64_t cb = blockIdx.x;
    const int64_t n_cb = gridDim.x;
    if (threadIdx.x < kColsPerBlock) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kBiasThreads / 32; ++w) t += red[w][threadIdx.x];
        partial[(((int64_t)s * n_cb + cb) * n_slabs + slab) * kColsPerBlock + threadIdx.x] = t;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(counters + s * n_cb + cb, 1u);
        is_last = (done == (unsigned int)n_slabs - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < kColsPerBlock) {
        const volatile float* pp = partial + ((int64_t)s * n_cb + cb) * n_slabs * kColsPerBlock + threadIdx.x;
        float t = 0.0f;
        for (int k = 0; k < n_slabs; ++k) t += pp[(int64_t)k * kColsPerBlock];
        const int64_t c = cb * kColsPerBlock + threadIdx.x;
        if (c < N) db[(int64_t)s * N + c] = t;
    }
    if (threadIdx.x == 0) counters[s * n_cb + cb] = 0u;
}

inline void bias_grid(int64_t S, int64_t M, int64_t N, int& n_cb, int& n_slabs, int64_t& rows_per_slab) {
    n_cb = (int)((N + kColsPerBlock - 1) / kColsPerBlock);
    const int64_t target = (int64_t)bf_num_sms() * 4;
    int64_t slabs = target / (S * n_cb);
    const int64_t max_slabs = (M + 63) / 64;
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    if (slabs > 65535) slabs = 65535;
    rows_per_slab = (M + slabs - 1) / slabs;
    n_slabs = (int)((M + rows_per_slab - 1) / rows_per_slab);
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
    int n_cb, n_slabs;
    int64_t rps;
    bias_grid(S < 1 ? 1 : S, M < 1 ? 1 : M, N < 1 ? 1 : N, n_cb, n_slabs, rps);
    // [counters: S*n_cb uint32, padded to 256 B][partials: S*n_cb*n_slabs*64 floats]
    const int64_t cnt = ((S * n_cb * 4 + 255) / 256) * 256;
    return cnt + S * n_cb * (int64_t)n_slabs * kColsPerBlock * 4;
}

extern "C" int bf_bias_grad(const void* gy, int32_t gy_dtype, float* db, int64_t S, int64_t M, int64_t N,
                            void* workspace, void* stream) {
    BF_CHECK_ARG(gy && db && workspace, "null pointer");
    BF_CHECK_ARG(gy_dtype == BF_F32 || gy_dtype == BF_BF16, "bad dtype");
    if (S == 0 || N == 0) return 0;
    int n_cb, n_slabs;
    int64_t rps;
    bias_grid(S, M < 1 ? 1 : M, N, n_cb, n_slabs, rps);
    const int64_t cnt = ((S * n_cb * 4 + 255) / 256) * 256;
    unsigned int* counters = reinterpret_cast<unsigned int*>(workspace);
    float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + cnt);
    dim3 grid((unsigned)n_cb, (unsigned)S, (unsigned)n_slabs);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (gy_dtype == BF_BF16)
        bias_grad_kernel<__nv_bfloat16><<<grid, kBiasThreads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(gy), db,
                                                                        partial, counters, M, N, rps);
    else
        bias_grad_kernel<float><<<grid, kBiasThreads, 0, st>>>(reinterpret_cast<const float*>(gy), db, partial,
                                                                counters, M, N, rps);
    BF_LAUNCH_OK();
    return 0;
}
