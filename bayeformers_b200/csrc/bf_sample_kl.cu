// Fused Philox sample + log q + log p (forward) and its stand-alone backward.
//
// Reference arithmetic being replaced (paths relative to /root/reference):
//   bayeformers/nn/parameters/gaussian.py:90-101    Gaussian.sample
//   bayeformers/nn/parameters/gaussian.py:103-116   Gaussian.log_prob (posterior q, MOPED prior p)
//   bayeformers/nn/parameters/gaussian.py:160-171   ScaledGaussianMixture.log_prob
//   bayeformers/nn/layers/linear.py:97-102          their call site in Linear.forward
//
// One pass over (mu, rho[, prior]) produces S weight samples and the 2*S
// scalar reductions.  HBM traffic per element: 8 B read (+8 B for a Gaussian
// prior, +0 when prior_mu aliases mu... still 4 B rho_p) and S*b_w written;
// eps lives only in registers.  Reductions are deterministic: per-thread
// sequential -> warp shuffle tree -> block tree -> fixed-order final pass by
// the last block to finish (no float atomics).
#include <cstdlib>

#include "bf_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxSC = 8;  // samples handled per launch (accumulators stay in registers)

struct SampleKlParams {
    const float* mu;
    const float* rho;
    const float* prior_mu;
    const float* prior_rho;
    const float* eps_in;  // [S_total * n] or null
    const float* sigma;   // multi-tensor path: cached softplus(rho), read instead of rho (or null)
    void* w_out;          // [S_total][w_stride] or null
    float* logq_out;      // [S_total]
    float* logp_out;
    float* partials;      // [2*SC][gridDim.x]
    unsigned int* counter;
    int64_t n;
    int64_t w_stride;
    int s0;  // first sample of this launch
    int accumulate;
    int64_t q0;  // first quad handled by this launch (generic kernel: tail of a vectorised run)
    uint32_t k0, k1, step, tensor_id;
    const uint32_t* step_ptr;  // optional device-resident offset added to `step`
    float prior_const_c, prior_const_iv;  // Gaussian prior with constant sigma (prior_rho == NULL)
    BfMixture mix;
};

template <typename T>
__device__ __forceinline__ void store4(T* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(a, b, c, d));
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 v;
    v.x = *reinterpret_cast<const uint32_t*>(&lo);
    v.y = *reinterpret_cast<const uint32_t*>(&hi);
    __stcs(reinterpret_cast<uint2*>(p), v);
}
template <typename T>
__device__ __forceinline__ void store1(T* p, float a);
template <>
__device__ __forceinline__ void store1<float>(float* p, float a) { *p = a; }
template <>
__device__ __forceinline__ void store1<__nv_bfloat16>(__nv_bfloat16* p, float a) { *p = __float2bfloat16_rn(a); }

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// one element, one sample: returns w and adds the two log-prob terms
template <int PRIOR>
__device__ __forceinline__ float element_terms(float mu, float sigma, float neg_c_minus_logsig, float inv_two_var,
                                               float eps, float pmu, float p_const, float p_inv_two_var,
                                               const BfMixture& mix, float& q_acc, float& p_acc) {
    // w = mu + eps*sigma with the reference's two roundings (no fma contraction)
    const float w = __fadd_rn(mu, __fmul_rn(eps, sigma));
    const float d = __fsub_rn(w, mu);
    q_acc += neg_c_minus_logsig - __fmul_rn(d, d) * inv_two_var;
    if (PRIOR == BF_PRIOR_MIXTURE) {
        p_acc += bf_mixture_logp(w, mix);
    } else if (PRIOR == BF_PRIOR_GAUSSIAN) {
        const float dp = __fsub_rn(w, pmu);
        p_acc += p_const - __fmul_rn(dp, dp) * p_inv_two_var;
    }
    return w;
}

// deterministic reduction tail shared by both forward kernels: per-thread values ->
// warp shuffle tree -> block tree -> per-block partial -> fixed-order final pass
// by the last block to arrive (no float atomics)
template <int SC>
__device__ __forceinline__ void block_finish(const SampleKlParams& p, float (&q_acc)[SC], float (&p_acc)[SC]) {
    // ---- block reduction of the 2*SC accumulators -------------------------
    __shared__ float red[2 * kMaxSC][kThreads / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < SC; ++s) {
        const float a = bf_warp_sum(q_acc[s]);
        const float b = bf_warp_sum(p_acc[s]);
        if (lane == 0) {
            red[2 * s][warp] = a;
            red[2 * s + 1][warp] = b;
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * SC) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) t += red[threadIdx.x][w];
        p.partials[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = t;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(p.counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- final fixed-order pass (one block): warp w handles value w, w+8, ..
    for (int v = warp; v < 2 * SC; v += kThreads / 32) {
        double t = 0.0;
        const volatile float* row = p.partials + (int64_t)v * gridDim.x;
        for (unsigned int b = lane; b < gridDim.x; b += 32) t += (double)row[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) {
            float* dst = ((v & 1) ? p.logp_out : p.logq_out) + p.s0 + (v >> 1);
            const float r = (float)t;
            *dst = p.accumulate ? (*dst + r) : r;
        }
    }
    if (threadIdx.x == 0) *p.counter = 0u;  // self-reset for the next launch
}

template <int PRIOR, typename WT, int SC, bool VEC>
__global__ void __launch_bounds__(kThreads) sample_kl_fwd_kernel(const SampleKlParams p) {
    float q_acc[SC], p_acc[SC];
#pragma unroll
    for (int s = 0; s < SC; ++s) q_acc[s] = p_acc[s] = 0.0f;

    WT* const w_out = reinterpret_cast<WT*>(p.w_out);
    const uint32_t step = p.step + (p.step_ptr ? __ldg(p.step_ptr) : 0u);
    const int64_t n = p.n;
    const int64_t nquad = (n + 3) >> 2;
    const int64_t stride = (int64_t)gridDim.x * kThreads;

    for (int64_t q = p.q0 + (int64_t)blockIdx.x * kThreads + threadIdx.x; q < nquad; q += stride) {
        const int64_t i0 = q << 2;
        const bool full = VEC && (i0 + 4 <= n);
        float mu[4], rho[4], pmu[4], prho[4];
        if (full) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p.mu + i0));
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.rho + i0));
            mu[0] = a.x, mu[1] = a.y, mu[2] = a.z, mu[3] = a.w;
            rho[0] = b.x, rho[1] = b.y, rho[2] = b.z, rho[3] = b.w;
            if (PRIOR == BF_PRIOR_GAUSSIAN) {
                const float4 c = __ldg(reinterpret_cast<const float4*>(p.prior_mu + i0));
                pmu[0] = c.x, pmu[1] = c.y, pmu[2] = c.z, pmu[3] = c.w;
                prho[0] = prho[1] = prho[2] = prho[3] = 0.0f;
                if (p.prior_rho != nullptr) {
                    const float4 d = __ldg(reinterpret_cast<const float4*>(p.prior_rho + i0));
                    prho[0] = d.x, prho[1] = d.y, prho[2] = d.z, prho[3] = d.w;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool ok = i0 + j < n;
                mu[j] = ok ? __ldg(p.mu + i0 + j) : 0.0f;
                rho[j] = ok ? __ldg(p.rho + i0 + j) : 0.0f;
                if (PRIOR == BF_PRIOR_GAUSSIAN) {
                    pmu[j] = ok ? __ldg(p.prior_mu + i0 + j) : 0.0f;
                    prho[j] = (ok && p.prior_rho) ? __ldg(p.prior_rho + i0 + j) : 0.0f;
                }
            }
        }
        // per-element quantities shared by the S samples
        float sigma[4], qc[4], qiv[4], pc[4], piv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sigma[j] = bf_softplus(rho[j]);
            qc[j] = -BF_LOG_SQRT_2PI - bf_log_sum_term(sigma[j]);
            qiv[j] = bf_rcp_approx(2.0f * __fmul_rn(sigma[j], sigma[j]));
            if (PRIOR == BF_PRIOR_GAUSSIAN) {
                if (p.prior_rho != nullptr) {
                    const float sp = bf_softplus(prho[j]);
                    pc[j] = -BF_LOG_SQRT_2PI - bf_log_sum_term(sp);
                    piv[j] = bf_rcp_approx(2.0f * __fmul_rn(sp, sp));
                } else {
                    pc[j] = p.prior_const_c, piv[j] = p.prior_const_iv;
                }
            } else {
                pc[j] = piv[j] = 0.0f;
                pmu[j] = 0.0f;
            }
        }
#pragma unroll
        for (int s = 0; s < SC; ++s) {
            const int sg = p.s0 + s;
            float e[4];
            if (p.eps_in != nullptr) {
                const float* ep = p.eps_in + (int64_t)sg * n + i0;
                if (full) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(ep));
                    e[0] = v.x, e[1] = v.y, e[2] = v.z, e[3] = v.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) e[j] = (i0 + j < n) ? __ldg(ep + j) : 0.0f;
                }
            } else {
                const float4 v = bf_eps_quad((uint32_t)q, (uint32_t)sg, p.tensor_id, step, p.k0, p.k1);
                e[0] = v.x, e[1] = v.y, e[2] = v.z, e[3] = v.w;
            }
            float w[4];
            if (full) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    w[j] = element_terms<PRIOR>(mu[j], sigma[j], qc[j], qiv[j], e[j], pmu[j], pc[j], piv[j], p.mix,
                                                q_acc[s], p_acc[s]);
                if (w_out != nullptr) store4<WT>(w_out + (int64_t)sg * p.w_stride + i0, w[0], w[1], w[2], w[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (i0 + j < n) {
                        w[j] = element_terms<PRIOR>(mu[j], sigma[j], qc[j], qiv[j], e[j], pmu[j], pc[j], piv[j],
                                                    p.mix, q_acc[s], p_acc[s]);
                        if (w_out != nullptr) store1<WT>(w_out + (int64_t)sg * p.w_stride + i0 + j, w[j]);
                    }
                }
            }
        }
    }

    block_finish<SC>(p, q_acc, p_acc);
}

// ---------------------------------------------------------------------------
// fast path: 16-byte aligned pointers, full quads only (the tail, if any, goes
// through the generic kernel above with q0 = n/4).  No bounds checks, injected
// eps resolved at compile time, two independent quads in flight per thread.
// ---------------------------------------------------------------------------
struct QuadShared {  // per-element quantities shared by the S samples of a quad
    float mu[4], sigma[4], qc[4], qiv[4], pmu[4], pc[4], piv[4];
};

struct RawQuad {  // the raw 16-byte loads of one quad, kept in flight across the compute of the previous quad
    float4 mu, rho, pmu, prho;
};

template <int PRIOR>
__device__ __forceinline__ void load_raw(const SampleKlParams& p, int64_t i0, RawQuad& R) {
    R.mu = __ldg(reinterpret_cast<const float4*>(p.mu + i0));
    R.rho = __ldg(reinterpret_cast<const float4*>((p.sigma ? p.sigma : p.rho) + i0));  // sigma cache: R.rho holds sigma
    if (PRIOR == BF_PRIOR_GAUSSIAN) {
        R.pmu = __ldg(reinterpret_cast<const float4*>(p.prior_mu + i0));
        if (p.prior_rho != nullptr) R.prho = __ldg(reinterpret_cast<const float4*>(p.prior_rho + i0));
    }
}

template <int PRIOR>
__device__ __forceinline__ void derive_quad(const SampleKlParams& p, const RawQuad& R, QuadShared& Q) {
    Q.mu[0] = R.mu.x, Q.mu[1] = R.mu.y, Q.mu[2] = R.mu.z, Q.mu[3] = R.mu.w;
    const float rho[4] = {R.rho.x, R.rho.y, R.rho.z, R.rho.w};
    float prho[4] = {0.f, 0.f, 0.f, 0.f};
    if (PRIOR == BF_PRIOR_GAUSSIAN) {
        Q.pmu[0] = R.pmu.x, Q.pmu[1] = R.pmu.y, Q.pmu[2] = R.pmu.z, Q.pmu[3] = R.pmu.w;
        if (p.prior_rho != nullptr) prho[0] = R.prho.x, prho[1] = R.prho.y, prho[2] = R.prho.z, prho[3] = R.prho.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        Q.sigma[j] = p.sigma ? rho[j] : bf_softplus(rho[j]);  // cached sigma (written by the optimizer) or rho
        Q.qc[j] = -BF_LOG_SQRT_2PI - bf_log_sum_term(Q.sigma[j]);
        Q.qiv[j] = bf_rcp_approx(2.0f * __fmul_rn(Q.sigma[j], Q.sigma[j]));
        if (PRIOR == BF_PRIOR_GAUSSIAN) {
            if (p.prior_rho != nullptr) {
                const float sp = bf_softplus(prho[j]);
                Q.pc[j] = -BF_LOG_SQRT_2PI - bf_log_sum_term(sp);
                Q.piv[j] = bf_rcp_approx(2.0f * __fmul_rn(sp, sp));
            } else {  // constant prior sigma (MOPED: rho_p == 1 everywhere), folded on the host
                Q.pc[j] = p.prior_const_c;
                Q.piv[j] = p.prior_const_iv;
            }
        } else {
            Q.pmu[j] = Q.pc[j] = Q.piv[j] = 0.0f;
        }
    }
}

// ---- packed fp32x2 form of the per-sample arithmetic (Gaussian prior or none) ---------------------------------
// The kernels are issue-bound (profiles/): FFMA2 / FMUL2 / FADD2 do two IEEE fp32 operations per issued instruction,
// and the sample-independent constants (-log sqrt(2 pi) - log sigma, prior ditto) are summed once per element instead
// of once per element-sample.  Every product / sum keeps the reference's separate roundings (w = mu + eps*sigma is
// mul.rn then add.rn), only the ORDER in which the log-prob terms are added differs (fp32 sums -> double at the end).
struct QuadPacked {
    bf_f2 mu[2], nmu[2], sigma[2], nqiv[2], npmu[2], npiv[2];
};

template <int PRIOR>
__device__ __forceinline__ void pack_quad(const QuadShared& Q, QuadPacked& P, float& qc_sum, float& pc_sum) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        P.mu[h] = bf_pack2(Q.mu[2 * h], Q.mu[2 * h + 1]);
        P.nmu[h] = bf_pack2(-Q.mu[2 * h], -Q.mu[2 * h + 1]);
        P.sigma[h] = bf_pack2(Q.sigma[2 * h], Q.sigma[2 * h + 1]);
        P.nqiv[h] = bf_pack2(-Q.qiv[2 * h], -Q.qiv[2 * h + 1]);
        if (PRIOR == BF_PRIOR_GAUSSIAN) {
            P.npmu[h] = bf_pack2(-Q.pmu[2 * h], -Q.pmu[2 * h + 1]);
            P.npiv[h] = bf_pack2(-Q.piv[2 * h], -Q.piv[2 * h + 1]);
        }
    }
    qc_sum += (Q.qc[0] + Q.qc[1]) + (Q.qc[2] + Q.qc[3]);
    if (PRIOR == BF_PRIOR_GAUSSIAN) pc_sum += (Q.pc[0] + Q.pc[1]) + (Q.pc[2] + Q.pc[3]);
}

// one quad, one sample: w[4]; q2 / p2 accumulate -(w-mu)^2/(2 sigma^2) and -(w-mu_p)^2/(2 sigma_p^2) as lane pairs
template <int PRIOR>
__device__ __forceinline__ void quad_sample_packed(const QuadPacked& P, const float (&e)[4], bf_f2& q2, bf_f2& p2,
                                                   float (&w)[4]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const bf_f2 w2 = bf_add2(P.mu[h], bf_mul2(bf_pack2(e[2 * h], e[2 * h + 1]), P.sigma[h]));
        const bf_f2 d2 = bf_add2(w2, P.nmu[h]);
        q2 = bf_fma2(bf_mul2(d2, d2), P.nqiv[h], q2);
        if (PRIOR == BF_PRIOR_GAUSSIAN) {
            const bf_f2 dp2 = bf_add2(w2, P.npmu[h]);
            p2 = bf_fma2(bf_mul2(dp2, dp2), P.npiv[h], p2);
        }
        bf_unpack2(w2, w[2 * h], w[2 * h + 1]);
    }
}

template <int PRIOR, typename WT, int SC, bool HAS_EPS, int QPT>
__global__ void __launch_bounds__(kThreads) sample_kl_fwd_fast_kernel(const SampleKlParams p) {
    float q_acc[SC], p_acc[SC];
#pragma unroll
    for (int s = 0; s < SC; ++s) q_acc[s] = p_acc[s] = 0.0f;
    WT* const w_out = reinterpret_cast<WT*>(p.w_out);
    const uint32_t step = p.step + (p.step_ptr ? __ldg(p.step_ptr) : 0u);
    const int64_t n = p.n;
    const int64_t nquad = n >> 2;  // full quads only
    const int64_t stride = (int64_t)gridDim.x * kThreads;

    int64_t qb = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    RawQuad raw[QPT];
#pragma unroll
    for (int u = 0; u < QPT; ++u)
        if (qb + u * stride < nquad) load_raw<PRIOR>(p, (qb + u * stride) << 2, raw[u]);

    bf_f2 q2[SC], p2[SC];  // packed partial sums (Gaussian prior / none: see quad_sample_packed)
    float qc_sum = 0.0f, pc_sum = 0.0f;
#pragma unroll
    for (int s = 0; s < SC; ++s) q2[s] = p2[s] = bf_splat2(0.0f);
    for (; qb < nquad; qb += stride * QPT) {
        QuadShared Q[QPT];
        QuadPacked QP[QPT];
        bool live[QPT];
#pragma unroll
        for (int u = 0; u < QPT; ++u) {
            live[u] = qb + u * stride < nquad;
            if (live[u]) {
                derive_quad<PRIOR>(p, raw[u], Q[u]);
                if (PRIOR != BF_PRIOR_MIXTURE) pack_quad<PRIOR>(Q[u], QP[u], qc_sum, pc_sum);
            }
        }
        // prefetch the next iteration's parameters: their latency hides behind this iteration's Philox work
#pragma unroll
        for (int u = 0; u < QPT; ++u) {
            const int64_t qn = qb + (QPT + u) * stride;
            if (qn < nquad) load_raw<PRIOR>(p, qn << 2, raw[u]);
        }
#pragma unroll
        for (int s = 0; s < SC; ++s) {
            const int sg = p.s0 + s;
#pragma unroll
            for (int u = 0; u < QPT; ++u) {
                if (!live[u]) continue;
                const int64_t q = qb + u * stride;
                const int64_t i0 = q << 2;
                float e[4];
                if (HAS_EPS) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(p.eps_in + (int64_t)sg * n + i0));
                    e[0] = v.x, e[1] = v.y, e[2] = v.z, e[3] = v.w;
                } else {
                    const float4 v = bf_eps_quad((uint32_t)q, (uint32_t)sg, p.tensor_id, step, p.k0, p.k1);
                    e[0] = v.x, e[1] = v.y, e[2] = v.z, e[3] = v.w;
                }
                float w[4];
                if (PRIOR != BF_PRIOR_MIXTURE) {
                    quad_sample_packed<PRIOR>(QP[u], e, q2[s], p2[s], w);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        w[j] = element_terms<PRIOR>(Q[u].mu[j], Q[u].sigma[j], Q[u].qc[j], Q[u].qiv[j], e[j], Q[u].pmu[j],
                                                    Q[u].pc[j], Q[u].piv[j], p.mix, q_acc[s], p_acc[s]);
                }
                if (w_out != nullptr) store4<WT>(w_out + (int64_t)sg * p.w_stride + i0, w[0], w[1], w[2], w[3]);
            }
        }
    }
    if (PRIOR != BF_PRIOR_MIXTURE) {
#pragma unroll
        for (int s = 0; s < SC; ++s) {
            float a, b;
            bf_unpack2(q2[s], a, b);
            q_acc[s] += (a + b) + qc_sum;
            if (PRIOR == BF_PRIOR_GAUSSIAN) {
                bf_unpack2(p2[s], a, b);
                p_acc[s] += (a + b) + pc_sum;
            }
        }
    }
    block_finish<SC>(p, q_acc, p_acc);
}

// ---------------------------------------------------------------------------
// stand-alone backward
// ---------------------------------------------------------------------------
struct SampleKlBwdParams {
    const void* grad_w;
    const float* mu;
    const float* rho;
    const float* prior_mu;
    const float* prior_rho;
    const float* g_logq;
    const float* g_logp;
    const float* eps_in;
    float* grad_mu;
    float* grad_rho;
    int64_t n;
    int64_t gw_stride;
    int S;
    int accumulate;
    uint32_t k0, k1, step, tensor_id;
    const uint32_t* step_ptr;
    float prior_const_ipv;  // 1/sigma_p^2 when the Gaussian prior has a constant sigma (prior_rho == NULL)
    BfMixture mix;
};

// 4 consecutive elements starting at i0 (vector access when `full`)
__device__ __forceinline__ void ld4(const float* p, int64_t i0, int64_t n, bool full, float (&v)[4]) {
    if (full) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p + i0));
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (i0 + j < n) ? __ldg(p + i0 + j) : 0.0f;
    }
}
__device__ __forceinline__ void ld4(const __nv_bfloat16* p, int64_t i0, int64_t n, bool full, float (&v)[4]) {
    if (full) {
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(p + i0));
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&a.x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&a.y);
        v[0] = __low2float(lo), v[1] = __high2float(lo), v[2] = __low2float(hi), v[3] = __high2float(hi);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (i0 + j < n) ? __bfloat162float(p[i0 + j]) : 0.0f;
    }
}
__device__ __forceinline__ void st4_acc(float* p, int64_t i0, int64_t n, bool full, const float (&v)[4], int acc) {
    if (full) {
        float4 o = make_float4(v[0], v[1], v[2], v[3]);
        if (acc) {
            const float4 a = *reinterpret_cast<const float4*>(p + i0);
            o.x += a.x, o.y += a.y, o.z += a.z, o.w += a.w;
        }
        *reinterpret_cast<float4*>(p + i0) = o;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i0 + j < n) p[i0 + j] = acc ? p[i0 + j] + v[j] : v[j];
    }
}

template <int PRIOR, typename GT, bool KL, bool VEC>
__global__ void __launch_bounds__(kThreads) sample_kl_bwd_kernel(const SampleKlBwdParams p) {
    const GT* const gw = reinterpret_cast<const GT*>(p.grad_w);
    const uint32_t step = p.step + (p.step_ptr ? __ldg(p.step_ptr) : 0u);
    const int64_t n = p.n;
    const int64_t nquad = (n + 3) >> 2;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x; q < nquad; q += stride) {
        const int64_t i0 = q << 2;
        const bool full = VEC && (i0 + 4 <= n);
        float rho[4], mu[4], pmu[4], sigma[4], inv_pvar[4], am[4], ar[4];
        ld4(p.rho, i0, n, full, rho);
#pragma unroll
        for (int j = 0; j < 4; ++j) am[j] = ar[j] = 0.0f, mu[j] = pmu[j] = sigma[j] = inv_pvar[j] = 0.0f;
        if (KL) {
            ld4(p.mu, i0, n, full, mu);
#pragma unroll
            for (int j = 0; j < 4; ++j) sigma[j] = bf_softplus(rho[j]);
            if (PRIOR == BF_PRIOR_GAUSSIAN) {
                ld4(p.prior_mu, i0, n, full, pmu);
                if (p.prior_rho != nullptr) {
                    float prho[4];
                    ld4(p.prior_rho, i0, n, full, prho);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float sp = bf_softplus(prho[j]);
                        inv_pvar[j] = 1.0f / (sp * sp);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) inv_pvar[j] = p.prior_const_ipv;
                }
            }
        }
        // software prefetch: the grad_w quad of sample s+1 is in flight while sample s is processed
        float g_next[4] = {0.f, 0.f, 0.f, 0.f};
        if (gw != nullptr) ld4(gw, i0, n, full, g_next);
        for (int s = 0; s < p.S; ++s) {
            float e[4], g[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] = g_next[j];
            if (gw != nullptr && s + 1 < p.S) ld4(gw + (int64_t)(s + 1) * p.gw_stride, i0, n, full, g_next);
            if (p.eps_in != nullptr) {
                ld4(p.eps_in + (int64_t)s * n, i0, n, full, e);
            } else {
                const float4 v = bf_eps_quad((uint32_t)q, (uint32_t)s, p.tensor_id, step, p.k0, p.k1);
                e[0] = v.x, e[1] = v.y, e[2] = v.z, e[3] = v.w;
            }
            float glq = 0.0f, glp = 0.0f;
            if (KL) {
                glq = p.g_logq ? __ldg(p.g_logq + s) : 0.0f;
                glp = p.g_logp ? __ldg(p.g_logp + s) : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float gj = g[j];
                if (KL) {
                    const float w = __fadd_rn(mu[j], __fmul_rn(e[j], sigma[j]));
                    float dp = 0.0f;
                    if (PRIOR == BF_PRIOR_MIXTURE) dp = bf_mixture_dlogp(w, p.mix);
                    if (PRIOR == BF_PRIOR_GAUSSIAN) dp = -(w - pmu[j]) * inv_pvar[j];
                    gj += glp * dp;             // chain through w (both the mu and the sigma*eps path)
                    ar[j] -= glq / sigma[j];    // d log q / d sigma = -1/sigma
                }
                am[j] += gj;
                ar[j] += gj * e[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) ar[j] *= bf_softplus_grad(rho[j]);
        st4_acc(p.grad_rho, i0, n, full, ar, p.accumulate);
        if (p.grad_mu != nullptr) st4_acc(p.grad_mu, i0, n, full, am, p.accumulate);
    }
}

__global__ void __launch_bounds__(kThreads) philox_normal_kernel(float* out, int64_t n, uint32_t k0, uint32_t k1,
                                                                 uint32_t step, uint32_t tensor_id, uint32_t sample) {
    const int64_t nquad = (n + 3) >> 2;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x; q < nquad; q += stride) {
        const float4 v = bf_eps_quad((uint32_t)q, sample, tensor_id, step, k0, k1);
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if ((q << 2) + j < n) out[(q << 2) + j] = e[j];
    }
}

// ---------------------------------------------------------------------------
// multi-tensor forward: every variational tensor of a model in ONE launch
// (SURVEY.md section 8f row 2).  The work list is cut into chunks of at most
// kChunkQuads quads of one tensor; blocks walk the chunk table grid-stride, so
// a BERT's 148 tensors (from 768-element biases to 3072x768 matrices) load the
// SMs evenly instead of paying 148 launches with ragged tails.  Reductions stay
// deterministic: per-chunk partial (block tree) -> per-slot sum in chunk order.
// ---------------------------------------------------------------------------
constexpr int kChunkQuads = 4096;

struct MultiParams {
    const bf_tensor_desc* descs;
    const int2* chunks;   // {tensor index, first quad}
    int n_chunks;
    int s0;               // first sample of this launch
    float* partials;      // [n_chunks][2*SC]
    uint32_t k0, k1, step;
    const uint32_t* step_ptr;
    char* w_base;  // when set, descs[].w_out are byte offsets from it
    int prefetch;  // 0: none, 1: prefetch.global.L1 of the next iteration's parameters, 2: prefetch.global.L2
};

template <int PRIOR, typename WT, int SC>
__device__ __forceinline__ void multi_chunk(const bf_tensor_desc& d, const MultiParams& mp, int64_t q_begin,
                                            uint32_t step, float (&q_acc)[SC], float (&p_acc)[SC]) {
    SampleKlParams p{};
    p.mu = d.mu, p.rho = d.rho, p.prior_mu = d.prior_mu, p.prior_rho = d.prior_rho;
    p.sigma = d.vec ? d.sigma : nullptr;  // the scalar path below reads rho itself
    p.mix.pi = d.pi;
    // w_out == (void*)-1: log-probs only (Embedding tables: rows are sampled on lookup, bf_embedding_fwd)
    const bool no_out = reinterpret_cast<intptr_t>(d.w_out) == (intptr_t)-1;
    WT* const w_out = no_out ? nullptr
                      : mp.w_base ? reinterpret_cast<WT*>(mp.w_base + reinterpret_cast<intptr_t>(d.w_out))
                                  : reinterpret_cast<WT*>(d.w_out);
    const int64_t n = d.n;
    const int64_t nquad = (n + 3) >> 2;
    int64_t q_end = q_begin + kChunkQuads;
    if (q_end > nquad) q_end = nquad;
    if (PRIOR == BF_PRIOR_GAUSSIAN && d.prior_rho == nullptr) {
        p.prior_const_c = -BF_LOG_SQRT_2PI - logf(d.sigma1);
        p.prior_const_iv = 1.0f / (2.0f * d.sigma1 * d.sigma1);
    }
    if (PRIOR == BF_PRIOR_MIXTURE) {  // same constants as bf_make_mixture (host), evaluated per block
        const float v1 = d.sigma1 * d.sigma1, v2 = d.sigma2 * d.sigma2;
        p.mix.one_minus_pi = 1.0f - d.pi;
        p.mix.inv_two_var1 = 1.0f / (2.0f * v1), p.mix.inv_two_var2 = 1.0f / (2.0f * v2);
        p.mix.log_s1 = logf(d.sigma1), p.mix.log_s2 = logf(d.sigma2);
        p.mix.inv_var1 = 1.0f / v1, p.mix.inv_var2 = 1.0f / v2;
    }
    if (d.vec == 0) {
        // ragged / unaligned tensors (biases of odd length, ...): scalar, bounds-checked
        for (int64_t q = q_begin + threadIdx.x; q < q_end; q += kThreads) {
            const int64_t i0 = q << 2;
            RawQuad R;
            float t[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool ok = i0 + j < n;
                t[0][j] = ok ? __ldg(d.mu + i0 + j) : 0.0f;
                t[1][j] = ok ? __ldg(d.rho + i0 + j) : 0.0f;
                t[2][j] = (ok && PRIOR == BF_PRIOR_GAUSSIAN) ? __ldg(d.prior_mu + i0 + j) : 0.0f;
                t[3][j] = (ok && PRIOR == BF_PRIOR_GAUSSIAN && d.prior_rho) ? __ldg(d.prior_rho + i0 + j) : 0.0f;
            }
            R.mu = make_float4(t[0][0], t[0][1], t[0][2], t[0][3]);
            R.rho = make_float4(t[1][0], t[1][1], t[1][2], t[1][3]);
            R.pmu = make_float4(t[2][0], t[2][1], t[2][2], t[2][3]);
            R.prho = make_float4(t[3][0], t[3][1], t[3][2], t[3][3]);
            QuadShared Q;
            derive_quad<PRIOR>(p, R, Q);
#pragma unroll
            for (int s = 0; s < SC; ++s) {
                const int sg = mp.s0 + s;
                const float4 v = bf_eps_quad((uint32_t)q, (uint32_t)sg, d.tensor_id, step + d.step, mp.k0, mp.k1);
                const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (i0 + j < n) {
                        const float w = element_terms<PRIOR>(Q.mu[j], Q.sigma[j], Q.qc[j], Q.qiv[j], e[j], Q.pmu[j], Q.pc[j],
                                                             Q.piv[j], p.mix, q_acc[s], p_acc[s]);
                        if (w_out != nullptr) store1<WT>(w_out + (int64_t)sg * d.w_stride + i0 + j, w);
                    }
                }
            }
        }
        return;
    }
    // vector path: 16-byte loads / stores, full quads only
    bf_f2 q2[SC], p2[SC];  // packed partial sums (see quad_sample_packed)
    float qc_sum = 0.0f, pc_sum = 0.0f;
#pragma unroll
    for (int s = 0; s < SC; ++s) q2[s] = p2[s] = bf_splat2(0.0f);
    for (int64_t q = q_begin + threadIdx.x; q < q_end; q += kThreads) {
        const int64_t i0 = q << 2;
        RawQuad R;
        load_raw<PRIOR>(p, i0, R);
        if (mp.prefetch && q + kThreads < q_end) {
            // the next iteration's parameter lines: without this every iteration exposes the full DRAM latency at
            // 16 resident warps per SM (ncu source page: 21 % of all stall samples on the first use of rho)
            const int64_t in = (q + kThreads) << 2;
            const float* const r_or_s = p.sigma ? p.sigma : d.rho;
            if (mp.prefetch == 1) {
                prefetch_l1(d.mu + in), prefetch_l1(r_or_s + in);
                if (PRIOR == BF_PRIOR_GAUSSIAN) prefetch_l1(d.prior_mu + in);
            } else {
                prefetch_l2(d.mu + in), prefetch_l2(r_or_s + in);
                if (PRIOR == BF_PRIOR_GAUSSIAN) prefetch_l2(d.prior_mu + in);
            }
        }
        if (PRIOR != BF_PRIOR_MIXTURE) {
            QuadPacked QP;
            {
                QuadShared Q;
                derive_quad<PRIOR>(p, R, Q);
                pack_quad<PRIOR>(Q, QP, qc_sum, pc_sum);
            }
#pragma unroll
            for (int s = 0; s < SC; ++s) {
                const int sg = mp.s0 + s;
                const float4 v = bf_eps_quad((uint32_t)q, (uint32_t)sg, d.tensor_id, step + d.step, mp.k0, mp.k1);
                const float e[4] = {v.x, v.y, v.z, v.w};
                float w[4];
                quad_sample_packed<PRIOR>(QP, e, q2[s], p2[s], w);
                if (w_out != nullptr) store4<WT>(w_out + (int64_t)sg * d.w_stride + i0, w[0], w[1], w[2], w[3]);
            }
        } else {
            QuadShared Q;
            derive_quad<PRIOR>(p, R, Q);
#pragma unroll
            for (int s = 0; s < SC; ++s) {
                const int sg = mp.s0 + s;
                const float4 v = bf_eps_quad((uint32_t)q, (uint32_t)sg, d.tensor_id, step + d.step, mp.k0, mp.k1);
                const float e[4] = {v.x, v.y, v.z, v.w};
                float w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    w[j] = element_terms<PRIOR>(Q.mu[j], Q.sigma[j], Q.qc[j], Q.qiv[j], e[j], Q.pmu[j], Q.pc[j], Q.piv[j],
                                                p.mix, q_acc[s], p_acc[s]);
                if (w_out != nullptr) store4<WT>(w_out + (int64_t)sg * d.w_stride + i0, w[0], w[1], w[2], w[3]);
            }
        }
    }
    if (PRIOR != BF_PRIOR_MIXTURE) {
#pragma unroll
        for (int s = 0; s < SC; ++s) {
            float a, b;
            bf_unpack2(q2[s], a, b);
            q_acc[s] += (a + b) + qc_sum;
            if (PRIOR == BF_PRIOR_GAUSSIAN) {
                bf_unpack2(p2[s], a, b);
                p_acc[s] += (a + b) + pc_sum;
            }
        }
    }
}

// One instantiation per prior kind (a block skips the chunks of other kinds): the scale-mixture arithmetic needs ~30
// registers more than the Gaussian one, and a kernel that inlines both runs every tensor at the larger footprint
// (96 registers = 2 blocks per SM; the Gaussian / no-prior instantiations fit 3).
template <int SC, int PRIOR>
__global__ void __launch_bounds__(kThreads, (PRIOR == BF_PRIOR_MIXTURE || SC > 4) ? 2 : 3) sample_kl_multi_kernel(const MultiParams mp) {
    __shared__ float red[2 * kMaxSC][kThreads / 32];
    const uint32_t step = mp.step + (mp.step_ptr ? __ldg(mp.step_ptr) : 0u);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = blockIdx.x; c < mp.n_chunks; c += gridDim.x) {
        const int2 ch = __ldg(mp.chunks + c);
        const bf_tensor_desc d = mp.descs[ch.x];
        if (d.prior_kind != PRIOR) continue;  // uniform per block
        float q_acc[SC], p_acc[SC];
#pragma unroll
        for (int s = 0; s < SC; ++s) q_acc[s] = p_acc[s] = 0.0f;
        const int64_t qb = (int64_t)ch.y;
        if (d.w_dtype == BF_BF16) multi_chunk<PRIOR, __nv_bfloat16, SC>(d, mp, qb, step, q_acc, p_acc);
        else multi_chunk<PRIOR, float, SC>(d, mp, qb, step, q_acc, p_acc);
#pragma unroll
        for (int s = 0; s < SC; ++s) {
            const float a = bf_warp_sum(q_acc[s]);
            const float b = bf_warp_sum(p_acc[s]);
            if (lane == 0) {
                red[2 * s][warp] = a;
                red[2 * s + 1][warp] = b;
            }
        }
        __syncthreads();
        if (threadIdx.x < 2 * SC) {
            float t = 0.0f;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) t += red[threadIdx.x][w];
            mp.partials[(int64_t)c * 2 * SC + threadIdx.x] = t;
        }
        __syncthreads();
    }
}

template <int PRIOR>
static void launch_multi(const MultiParams& mp, int sc, int grid, cudaStream_t st) {
    switch (sc) {
        case 1: sample_kl_multi_kernel<1, PRIOR><<<grid, kThreads, 0, st>>>(mp); break;
        case 2: sample_kl_multi_kernel<2, PRIOR><<<grid, kThreads, 0, st>>>(mp); break;
        case 4: sample_kl_multi_kernel<4, PRIOR><<<grid, kThreads, 0, st>>>(mp); break;
        default: sample_kl_multi_kernel<8, PRIOR><<<grid, kThreads, 0, st>>>(mp); break;
    }
}

// one block per output slot: sum the chunk partials of that slot in chunk order
__global__ void __launch_bounds__(256) sample_kl_multi_finish_kernel(const float* __restrict__ partials,
                                                                    const int2* __restrict__ slot_ranges, int SC, int s0,
                                                                    int S, float* __restrict__ logq,
                                                                    float* __restrict__ logp) {
    const int slot = blockIdx.x;
    const int2 r = __ldg(slot_ranges + slot);  // chunks [r.x, r.y)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int v = warp; v < 2 * SC; v += 8) {  // one warp per value; fixed lane-strided order, then a shuffle tree
        double t = 0.0;
#pragma unroll 4
        for (int c = r.x + lane; c < r.y; c += 32) t += (double)partials[(int64_t)c * 2 * SC + v];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) ((v & 1) ? logp : logq)[(int64_t)slot * S + s0 + (v >> 1)] = (float)t;
    }
}

inline int grid_for(int64_t nquad, int blocks_per_sm) {
    const int64_t want = (nquad + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)bf_num_sms() * blocks_per_sm;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

constexpr int kFwdBlocksPerSm = 8;

// ---- generic (scalar, bounds-checked) kernel: one sample per launch --------------
template <int PRIOR>
void launch_generic(const SampleKlParams& p, int grid, int w_dtype, cudaStream_t st) {
    if (w_dtype == BF_BF16)
        sample_kl_fwd_kernel<PRIOR, __nv_bfloat16, 1, false><<<grid, kThreads, 0, st>>>(p);
    else
        sample_kl_fwd_kernel<PRIOR, float, 1, false><<<grid, kThreads, 0, st>>>(p);
}

// ---- fast kernel: SC samples per launch, QPT quads in flight per thread ----------
template <int PRIOR, typename WT, bool HAS_EPS>
int launch_fast_sc(const SampleKlParams& p, int sc, int grid, cudaStream_t st) {
    switch (sc) {
        case 1: sample_kl_fwd_fast_kernel<PRIOR, WT, 1, HAS_EPS, 2><<<grid, kThreads, 0, st>>>(p); break;
        case 2: sample_kl_fwd_fast_kernel<PRIOR, WT, 2, HAS_EPS, 2><<<grid, kThreads, 0, st>>>(p); break;
        case 4: sample_kl_fwd_fast_kernel<PRIOR, WT, 4, HAS_EPS, 1><<<grid, kThreads, 0, st>>>(p); break;
        case 8: sample_kl_fwd_fast_kernel<PRIOR, WT, 8, HAS_EPS, 1><<<grid, kThreads, 0, st>>>(p); break;
        default: return BF_ERR_BAD_ARG;
    }
    return 0;
}

template <int PRIOR>
int launch_fast(const SampleKlParams& p, int sc, int grid, int w_dtype, cudaStream_t st) {
    const bool eps = p.eps_in != nullptr;
    if (w_dtype == BF_BF16)
        return eps ? launch_fast_sc<PRIOR, __nv_bfloat16, true>(p, sc, grid, st)
                   : launch_fast_sc<PRIOR, __nv_bfloat16, false>(p, sc, grid, st);
    return eps ? launch_fast_sc<PRIOR, float, true>(p, sc, grid, st)
               : launch_fast_sc<PRIOR, float, false>(p, sc, grid, st);
}

inline bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; }

}  // namespace

extern "C" int64_t bf_sample_kl_workspace_bytes(int64_t n, int32_t S) {
    (void)n;
    (void)S;
    // [counter, padded to 256 B][2*kMaxSC rows of per-block partials]
    const int64_t max_grid = (int64_t)bf_num_sms() * kFwdBlocksPerSm;
    return 256 + (int64_t)2 * kMaxSC * max_grid * (int64_t)sizeof(float);
}

extern "C" int bf_sample_kl_fwd(const float* mu, const float* rho, int32_t prior_kind, const float* prior_mu,
                                const float* prior_rho, float pi, float sigma1, float sigma2, int64_t n, int32_t S,
                                uint64_t seed, uint32_t step, uint32_t tensor_id, const float* eps_in, void* w_out,
                                int32_t w_dtype, int64_t w_stride, float* logq_out, float* logp_out,
                                int32_t accumulate, void* workspace, void* stream) {
    BF_CHECK_ARG(logq_out && logp_out && workspace, "null pointer");
    BF_CHECK_ARG(n >= 0 && S >= 1, "bad n or S");
    BF_CHECK_ARG(n == 0 || (mu && rho), "null mu/rho");
    BF_CHECK_ARG(prior_kind == BF_PRIOR_MIXTURE || prior_kind == BF_PRIOR_GAUSSIAN || prior_kind == BF_PRIOR_NONE,
                 "bad prior_kind");
    BF_CHECK_ARG(prior_kind != BF_PRIOR_GAUSSIAN || n == 0 || prior_mu, "gaussian prior needs prior_mu");
    BF_CHECK_ARG(w_dtype == BF_F32 || w_dtype == BF_BF16, "bad w_dtype");
    BF_CHECK_ARG(w_out == nullptr || w_stride >= n, "w_stride < n");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    SampleKlParams p{};
    p.mu = mu, p.rho = rho, p.prior_mu = prior_mu, p.prior_rho = prior_rho, p.eps_in = eps_in;
    p.w_out = w_out, p.logq_out = logq_out, p.logp_out = logp_out;
    p.counter = reinterpret_cast<unsigned int*>(workspace);
    p.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
    p.n = n, p.w_stride = w_stride, p.accumulate = accumulate;
    p.k0 = (uint32_t)(seed & 0xffffffffu), p.k1 = (uint32_t)(seed >> 32);
    p.step = step, p.tensor_id = tensor_id, p.step_ptr = bf_step_counter();
    p.mix = bf_make_mixture(pi, sigma1, sigma2);

    if (prior_kind == BF_PRIOR_GAUSSIAN && prior_rho == nullptr) {
        // constant prior sigma passed in `sigma1` (see header): fold its terms on the host
        p.prior_const_c = -BF_LOG_SQRT_2PI - logf(sigma1);
        p.prior_const_iv = 1.0f / (2.0f * sigma1 * sigma1);
    }
    const int esz = w_dtype == BF_BF16 ? 2 : 4;
    bool vec = aligned16(mu) && aligned16(rho) && (eps_in == nullptr || (aligned16(eps_in) && n % 4 == 0));
    if (prior_kind == BF_PRIOR_GAUSSIAN) vec = vec && aligned16(prior_mu) && (!prior_rho || aligned16(prior_rho));
    if (w_out != nullptr) vec = vec && aligned16(w_out) && ((w_stride * esz) % 16 == 0);
    const int64_t nq_fast = vec ? (n >> 2) : 0;  // quads the fast kernel takes; the rest is the generic tail

#define BF_BY_PRIOR(CALL)                                                        \
    (prior_kind == BF_PRIOR_MIXTURE ? CALL<BF_PRIOR_MIXTURE> :                    \
     prior_kind == BF_PRIOR_GAUSSIAN ? CALL<BF_PRIOR_GAUSSIAN> : CALL<BF_PRIOR_NONE>)

    if (nq_fast > 0) {
        const int grid = grid_for(nq_fast, kFwdBlocksPerSm);
        for (int s0 = 0; s0 < S;) {
            int sc = 1;
            while (sc * 2 <= kMaxSC && s0 + sc * 2 <= S) sc *= 2;
            p.s0 = s0, p.q0 = 0, p.accumulate = accumulate;
            int rc = prior_kind == BF_PRIOR_MIXTURE    ? launch_fast<BF_PRIOR_MIXTURE>(p, sc, grid, w_dtype, st)
                     : prior_kind == BF_PRIOR_GAUSSIAN ? launch_fast<BF_PRIOR_GAUSSIAN>(p, sc, grid, w_dtype, st)
                                                       : launch_fast<BF_PRIOR_NONE>(p, sc, grid, w_dtype, st);
            if (rc) {
                bf_set_error("bf_sample_kl_fwd: bad sample chunk");
                return rc;
            }
            BF_LAUNCH_OK();
            s0 += sc;
        }
    }
    if ((nq_fast << 2) < n || n == 0) {
        const int64_t tail_quads = ((n + 3) >> 2) - nq_fast;
        const int grid = grid_for(tail_quads, kFwdBlocksPerSm);
        for (int s0 = 0; s0 < S; ++s0) {
            p.s0 = s0, p.q0 = nq_fast, p.accumulate = (nq_fast > 0) ? 1 : accumulate;
            if (prior_kind == BF_PRIOR_MIXTURE) launch_generic<BF_PRIOR_MIXTURE>(p, grid, w_dtype, st);
            else if (prior_kind == BF_PRIOR_GAUSSIAN) launch_generic<BF_PRIOR_GAUSSIAN>(p, grid, w_dtype, st);
            else launch_generic<BF_PRIOR_NONE>(p, grid, w_dtype, st);
            BF_LAUNCH_OK();
        }
    }
#undef BF_BY_PRIOR
    return 0;
}

extern "C" int bf_sample_kl_bwd(const void* grad_w, int32_t gw_dtype, int64_t gw_stride, const float* mu,
                                const float* rho, int32_t prior_kind, const float* prior_mu, const float* prior_rho,
                                float pi, float sigma1, float sigma2, const float* g_logq, const float* g_logp,
                                int64_t n, int32_t S, uint64_t seed, uint32_t step, uint32_t tensor_id,
                                const float* eps_in, float* grad_mu, float* grad_rho, int32_t accumulate,
                                void* stream) {
    BF_CHECK_ARG(rho && grad_rho, "null pointer");
    BF_CHECK_ARG(n >= 0 && S >= 1, "bad n or S");
    BF_CHECK_ARG(gw_dtype == BF_F32 || gw_dtype == BF_BF16, "bad gw_dtype");
    const bool kl = (g_logq != nullptr) || (g_logp != nullptr);
    BF_CHECK_ARG(!kl || mu, "KL gradient needs mu");
    BF_CHECK_ARG(!kl || prior_kind != BF_PRIOR_GAUSSIAN || prior_mu, "gaussian prior needs prior_mu");
    BF_CHECK_ARG(grad_w || kl, "nothing to do");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    SampleKlBwdParams p{};
    p.grad_w = grad_w, p.mu = mu, p.rho = rho, p.prior_mu = prior_mu, p.prior_rho = prior_rho;
    p.g_logq = g_logq, p.g_logp = g_logp, p.eps_in = eps_in, p.grad_mu = grad_mu, p.grad_rho = grad_rho;
    p.n = n, p.gw_stride = gw_stride, p.S = S, p.accumulate = accumulate;
    p.k0 = (uint32_t)(seed & 0xffffffffu), p.k1 = (uint32_t)(seed >> 32);
    p.step = step, p.tensor_id = tensor_id, p.step_ptr = bf_step_counter();
    p.mix = bf_make_mixture(pi, sigma1, sigma2);
    p.prior_const_ipv = 1.0f / (sigma1 * sigma1);
    const int grid = grid_for((n + 3) >> 2, 8);

    bool vec = aligned16(rho) && aligned16(grad_rho) && (!grad_mu || aligned16(grad_mu)) &&
               (!eps_in || (aligned16(eps_in) && n % 4 == 0));
    if (kl) vec = vec && aligned16(mu) && (prior_kind != BF_PRIOR_GAUSSIAN || (aligned16(prior_mu) && (!prior_rho || aligned16(prior_rho))));
    if (grad_w) vec = vec && ((reinterpret_cast<uintptr_t>(grad_w) & 15u) == 0) &&
                      ((gw_stride * (gw_dtype == BF_BF16 ? 2 : 4)) % 16 == 0);
#define BF_BWD(PRIOR, GT, KLF)                                                       \
    do {                                                                             \
        if (vec) sample_kl_bwd_kernel<PRIOR, GT, KLF, true><<<grid, kThreads, 0, st>>>(p);  \
        else sample_kl_bwd_kernel<PRIOR, GT, KLF, false><<<grid, kThreads, 0, st>>>(p);     \
    } while (0)
    const bool bf = gw_dtype == BF_BF16;
    if (!kl) {
        if (bf) BF_BWD(BF_PRIOR_NONE, __nv_bfloat16, false);
        else BF_BWD(BF_PRIOR_NONE, float, false);
    } else if (prior_kind == BF_PRIOR_MIXTURE) {
        if (bf) BF_BWD(BF_PRIOR_MIXTURE, __nv_bfloat16, true);
        else BF_BWD(BF_PRIOR_MIXTURE, float, true);
    } else if (prior_kind == BF_PRIOR_GAUSSIAN) {
        if (bf) BF_BWD(BF_PRIOR_GAUSSIAN, __nv_bfloat16, true);
        else BF_BWD(BF_PRIOR_GAUSSIAN, float, true);
    } else {
        if (bf) BF_BWD(BF_PRIOR_NONE, __nv_bfloat16, true);
        else BF_BWD(BF_PRIOR_NONE, float, true);
    }
#undef BF_BWD
    BF_LAUNCH_OK();
    return 0;
}

extern "C" int bf_philox_normal(float* out, int64_t n, uint64_t seed, uint32_t step, uint32_t tensor_id,
                                uint32_t sample_id, void* stream) {
    BF_CHECK_ARG(out && n >= 0, "bad args");
    if (n == 0) return 0;
    const int grid = grid_for((n + 3) >> 2, 8);
    philox_normal_kernel<<<grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        out, n, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), step, tensor_id, sample_id);
    BF_LAUNCH_OK();
    return 0;
}

// KL-only variational backward, accumulated on top of existing gradients (used by the
// fused weight-gradient path, whose tensor-core epilogue handles the data term only).
int bf_sample_kl_bwd_impl_kl_only(const float* mu, const float* rho, int32_t prior_kind, const float* prior_mu,
                                  const float* prior_rho, float pi, float sigma1, float sigma2, const float* g_logq,
                                  const float* g_logp, int64_t n, int32_t S, uint64_t seed, uint32_t step,
                                  uint32_t tensor_id, const float* eps_in, float* grad_mu, float* grad_rho,
                                  cudaStream_t st) {
    return bf_sample_kl_bwd(nullptr, BF_F32, n, mu, rho, prior_kind, prior_mu, prior_rho, pi, sigma1, sigma2, g_logq,
                            g_logp, n, S, seed, step, tensor_id, eps_in, grad_mu, grad_rho, 1, st);
}

namespace {
__global__ void __launch_bounds__(kThreads) softplus_fwd_kernel(const float* __restrict__ rho, float* __restrict__ sigma,
                                                                int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        sigma[i] = bf_softplus(__ldg(rho + i));
}
}  // namespace

extern "C" int bf_softplus_fwd(const float* rho, float* sigma, int64_t n, void* stream) {
    BF_CHECK_ARG(n >= 0, "bad n");
    if (n == 0) return 0;
    BF_CHECK_ARG(rho && sigma, "null pointer");
    softplus_fwd_kernel<<<grid_for((n + 3) >> 2, 8), kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rho, sigma, n);
    BF_LAUNCH_OK();
    return 0;
}

// ---- multi-tensor entry points --------------------------------------------------------
extern "C" int32_t bf_sample_kl_multi_chunk_quads(void) { return kChunkQuads; }

extern "C" int64_t bf_sample_kl_multi_workspace_bytes(int64_t n_chunks) {
    return (n_chunks < 1 ? 1 : n_chunks) * 2 * kMaxSC * (int64_t)sizeof(float);
}

extern "C" int bf_sample_kl_fwd_multi(const bf_tensor_desc* descs, const int32_t* chunks, int32_t n_chunks,
                                      const int32_t* slot_ranges, int32_t n_slots, int32_t prior_mask, int32_t S,
                                      uint64_t seed, uint32_t step, float* logq_out, float* logp_out, void* workspace,
                                      void* w_base, void* stream) {
    BF_CHECK_ARG(descs && chunks && slot_ranges && logq_out && logp_out && workspace, "null pointer");
    BF_CHECK_ARG(n_chunks >= 1 && n_slots >= 1 && S >= 1, "bad counts");
    BF_CHECK_ARG(prior_mask > 0 && prior_mask < 8, "prior_mask: bit k set when a descriptor has prior_kind == k");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    MultiParams mp{};
    mp.descs = descs, mp.chunks = reinterpret_cast<const int2*>(chunks), mp.n_chunks = n_chunks;
    mp.partials = reinterpret_cast<float*>(workspace);
    mp.k0 = (uint32_t)(seed & 0xffffffffu), mp.k1 = (uint32_t)(seed >> 32), mp.step = step;
    mp.step_ptr = bf_step_counter();
    mp.w_base = reinterpret_cast<char*>(w_base);
    mp.prefetch = bf_option(BF_OPT_SK_PREFETCH);  // A/B switch: 0 none, 1 L1 (default), 2 L2
    const int64_t cap = (int64_t)bf_num_sms() * kFwdBlocksPerSm;
    const int grid = (int)(n_chunks < cap ? n_chunks : cap);
    for (int s0 = 0; s0 < S;) {
        int sc = 1;
        while (sc * 2 <= kMaxSC && s0 + sc * 2 <= S) sc *= 2;
        mp.s0 = s0;
        if (prior_mask & (1 << BF_PRIOR_MIXTURE)) launch_multi<BF_PRIOR_MIXTURE>(mp, sc, grid, st);
        if (prior_mask & (1 << BF_PRIOR_GAUSSIAN)) launch_multi<BF_PRIOR_GAUSSIAN>(mp, sc, grid, st);
        if (prior_mask & (1 << BF_PRIOR_NONE)) launch_multi<BF_PRIOR_NONE>(mp, sc, grid, st);
        BF_LAUNCH_OK();
        sample_kl_multi_finish_kernel<<<n_slots, 256, 0, st>>>(mp.partials, reinterpret_cast<const int2*>(slot_ranges), sc,
                                                             s0, S, logq_out, logp_out);
        BF_LAUNCH_OK();
        s0 += sc;
    }
    return 0;
}
