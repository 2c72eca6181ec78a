// Shared device/host helpers for the bayeformers_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "../../include/bayeformers_b200.h"

// ---------------------------------------------------------------------------
// error plumbing (C ABI: return codes + thread-local text)
// ---------------------------------------------------------------------------
void bf_set_error(const std::string& msg);

#define BF_CHECK_ARG(cond, msg)                                        \
    do {                                                               \
        if (!(cond)) {                                                 \
            bf_set_error(std::string(__func__) + ": " + (msg));        \
            return BF_ERR_BAD_ARG;                                     \
        }                                                              \
    } while (0)

#define BF_CUDA_OK(expr)                                                                  \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            bf_set_error(std::string(__func__) + ": " #expr ": " + cudaGetErrorString(_e)); \
            return (int)_e;                                                               \
        }                                                                                 \
    } while (0)

#define BF_LAUNCH_OK()                                                                   \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            bf_set_error(std::string(__func__) + ": launch: " + cudaGetErrorString(_e)); \
            return (int)_e;                                                              \
        }                                                                                \
    } while (0)

int bf_num_sms();
// device-resident step counter (CUDA-graph replays must not redraw the same eps): see bf_set_step_counter
const uint32_t* bf_step_counter();
// tuning switch `id` (BF_OPT_*), see bf_set_option
int bf_option(int id);

// ---------------------------------------------------------------------------
// constants
// ---------------------------------------------------------------------------
// log(sqrt(2*pi)) as the reference folds it into fp32 arithmetic
// (numpy float64 constant, rounded to fp32 by the tensor op: gaussian.py:113)
#define BF_LOG_SQRT_2PI 0.91893853320467274178f

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Contract: oracle/philox_oracle.py
// ---------------------------------------------------------------------------
#define BF_PHILOX_M0 0xD2511F53u
#define BF_PHILOX_M1 0xCD9E8D57u
#define BF_PHILOX_W0 0x9E3779B9u
#define BF_PHILOX_W1 0xBB67AE85u

__device__ __forceinline__ uint4 bf_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                  uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(BF_PHILOX_M0, c0), lo0 = BF_PHILOX_M0 * c0;
        const uint32_t hi1 = __umulhi(BF_PHILOX_M1, c2), lo1 = BF_PHILOX_M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += BF_PHILOX_W0;
        k1 += BF_PHILOX_W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// MUFU-only sqrt / reciprocal (no IEEE slow-path branches in the hot loops)
__device__ __forceinline__ float bf_sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float bf_rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// two uint32 -> two N(0,1).  u in (0,1], angle in (-pi, pi] so the MUFU
// sin/cos approximations stay in their accurate range.
__device__ __forceinline__ void bf_box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    const float u = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // a*2^-32 + 2^-33
    const float f = fmaf((float)b, 4.6566128730773926e-10f, 2.3283064365386963e-10f - 1.0f);  // b*2^-31 + 2^-32 - 1
    // -2 ln u = -2 ln2 * log2 u
    const float radius = bf_sqrt_approx(-1.3862943611198906f * __log2f(u));
    float s, c;
    __sincosf(3.14159265358979323846f * f, &s, &c);
    n0 = radius * c;
    n1 = radius * s;
}

// eps for the 4 elements of quad `q` of (tensor_id, sample, step) under seed
__device__ __forceinline__ float4 bf_eps_quad(uint32_t q, uint32_t sample, uint32_t tensor_id, uint32_t step,
                                              uint32_t k0, uint32_t k1) {
    const uint4 r = bf_philox4x32_10(q, sample, tensor_id, step, k0, k1);
    float4 e;
    bf_box_muller(r.x, r.y, e.x, e.y);
    bf_box_muller(r.z, r.w, e.z, e.w);
    return e;
}

// ---------------------------------------------------------------------------
// elementwise math of the variational parameters
// ---------------------------------------------------------------------------
// ---- packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per issued instruction) ----
typedef unsigned long long bf_f2;  // two floats in one 64-bit register pair
__device__ __forceinline__ bf_f2 bf_pack2(float a, float b) {
    bf_f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void bf_unpack2(bf_f2 v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ bf_f2 bf_fma2(bf_f2 a, bf_f2 b, bf_f2 c) {
    bf_f2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ bf_f2 bf_mul2(bf_f2 a, bf_f2 b) {
    bf_f2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ bf_f2 bf_add2(bf_f2 a, bf_f2 b) {
    bf_f2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ bf_f2 bf_splat2(float c) { return bf_pack2(c, c); }

// ---- MUFU-based transcendental helpers ------------------------------------------------------
// The sample+KL kernels are issue-bound, not HBM-bound (profiles/README.md): libdevice's expf / log1pf /
// logf cost ~80 instructions per element with their slow-path branches.  These keep ~1e-6 relative
// accuracy (the parity bar is 1e-5 on sums, 1e-6 norm-wise on w) in ~25.
__device__ __forceinline__ float bf_ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float bf_lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// exp(x) for x <= ~20: x*log2(e) split into a rounded product and its fma residual, so the error does
// not grow with |x| the way __expf's does
__device__ __forceinline__ float bf_exp(float x) {
    const float t = x * 1.4426950408889634f;
    float r = fmaf(x, 1.4426950408889634f, -t);
    r = fmaf(x, 1.925963033500e-8f, r);  // low part of log2(e)
    const float z = bf_ex2_approx(t);
    return fmaf(z, r * 0.6931471805599453f, z);
}
// log(1 + z), z >= 0.  Small z: 2 atanh(s), s = z / (2 + z), odd series (s <= 0.2, truncation < 1e-8 relative);
// otherwise lg2.approx(1 + z) whose 2^-22 absolute error is relative to a result >= 0.4.
__device__ __forceinline__ float bf_log1p_pos(float z) {
    if (z < 0.5f) {
        const float s = z * bf_rcp_approx(2.0f + z);
        const float s2 = s * s;
        float p = fmaf(s2, 1.0f / 9.0f, 1.0f / 7.0f);
        p = fmaf(p, s2, 0.2f);
        p = fmaf(p, s2, 1.0f / 3.0f);
        p = fmaf(p, s2, 1.0f);
        return 2.0f * s * p;
    }
    return bf_lg2_approx(1.0f + z) * 0.6931471805599453f;
}
// log(x) for the per-element -log(sigma) term of a SUM of log-probs: absolute error ~2e-7
__device__ __forceinline__ float bf_log_sum_term(float x) { return bf_lg2_approx(x) * 0.6931471805599453f; }

// sigma = softplus(rho), beta=1, threshold=20 (torch F.softplus; gaussian.py:88)
__device__ __forceinline__ float bf_softplus(float rho) { return rho > 20.0f ? rho : bf_log1p_pos(bf_exp(rho)); }

// d softplus / d rho as torch's softplus_backward computes it: z/(z+1), z = exp(rho)
__device__ __forceinline__ float bf_softplus_grad(float rho) {
    if (rho > 20.0f) return 1.0f;
    const float z = expf(rho);
    return z / (z + 1.0f);
}

struct BfMixture {
    // fp32 constants of ScaledGaussianMixture.log_prob (gaussian.py:169-171 on
    // torch.distributions.Normal.log_prob): N_i = -(w^2)/(2 var_i) - log s_i - c
    float pi, one_minus_pi;
    float inv_two_var1, inv_two_var2;
    float log_s1, log_s2;
    float inv_var1, inv_var2;  // for d log p / dw
};

__host__ inline BfMixture bf_make_mixture(float pi, float s1, float s2) {
    BfMixture m;
    m.pi = pi;
    m.one_minus_pi = 1.0f - pi;
    const float v1 = s1 * s1, v2 = s2 * s2;
    m.inv_two_var1 = 1.0f / (2.0f * v1);
    m.inv_two_var2 = 1.0f / (2.0f * v2);
    m.log_s1 = logf(s1);
    m.log_s2 = logf(s2);
    m.inv_var1 = 1.0f / v1;
    m.inv_var2 = 1.0f / v2;
    return m;
}

// log(pi*exp(N1) + (1-pi)*exp(N2)) -- explicit exp-then-log like the reference,
// so the fp32 underflow behaviour (component 2 vanishing, -inf for |w| >~ 13) is kept.
__device__ __forceinline__ float bf_mixture_logp(float w, const BfMixture& m) {
    const float w2 = w * w;
    const float n1 = -(w2 * m.inv_two_var1) - m.log_s1 - BF_LOG_SQRT_2PI;
    const float n2 = -(w2 * m.inv_two_var2) - m.log_s2 - BF_LOG_SQRT_2PI;
    if (n1 > -80.0f) {
        // common case (|w| < ~12 sigma1): component 1 is a normal fp32 number, MUFU ex2 / lg2 with an fma-corrected
        // argument are accurate to ~1e-7 here.  Component 2 may flush to zero where expf would still return a
        // denormal; it is then < 1e-30 of component 1 and changes nothing in fp32.
        return bf_log_sum_term(m.pi * bf_exp(n1) + m.one_minus_pi * bf_exp(n2));
    }
    // far tail: libdevice, so the reference's denormal / -inf behaviour (quirk Q10) is reproduced exactly
    return logf(m.pi * expf(n1) + m.one_minus_pi * expf(n2));
}

// d log p / dw of the mixture, in the max-shifted (overflow-safe) form
__device__ __forceinline__ float bf_mixture_dlogp(float w, const BfMixture& m) {
    const float w2 = w * w;
    const float l1 = __logf(m.pi) - m.log_s1 - w2 * m.inv_two_var1;
    const float l2 = __logf(m.one_minus_pi) - m.log_s2 - w2 * m.inv_two_var2;
    const float mx = fmaxf(l1, l2);
    const float a1 = __expf(l1 - mx), a2 = __expf(l2 - mx);
    return -w * (a1 * m.inv_var1 + a2 * m.inv_var2) / (a1 + a2);
}

// ---------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------
__device__ __forceinline__ float bf_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float bf_ld_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float bf_ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }
