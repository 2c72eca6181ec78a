// Fused  y = LayerNorm(dropout(h) + r)  forward / backward for the S folded samples (sm_100a, HBM-bound).
//
// This is the code either side of a Bayesian Linear in a transformer "output" block
// (HuggingFace BertSelfOutput / BertOutput: dense -> dropout -> LayerNorm(h + input)); SURVEY.md
// section 8f lists the callers around the path as the next rows once the path itself meets the bar.
// The LayerNorm is the S-sample LayerNorm of row A10 (per-sample gamma_s / beta_s drawn from
// bayeformers/nn/parameters/gaussian.py:90-101 semantics) or, with affine_stride == 0, the host
// model's frequentist LayerNorm.
//
// Unfused, the block costs 8 passes over [S*M, H] in the forward (dropout 3 incl. the byte mask,
// residual add 3, LayerNorm 2) and 6 + a bias-gradient pass in the backward.  Fused:
//   fwd : read h, r            -> write z = dropout(h) + r (kept for backward), y           (4 passes)
//   bwd : read gy, z           -> write dz (gradient of r), dh = dz * mask / (1 - p),
//         plus dgamma, dbeta and the column sums of dh per sample (the bias gradient of the
//         Linear that produced h), all in the same pass                                     (4 passes)
// The dropout mask is never stored: like eps it is a pure function of a Philox4x32-10 counter
// (seed, site, step, element / 8) -- 16 random bits per element, keep iff u16 >= round(p * 65536) --
// and is regenerated in the backward.
//
// One warp owns one row, held in registers as 16-byte chunks (H = 256*C, lane l holds elements
// [c*256 + l*8, +8)); statistics are two-pass fp32 on the stored (rounded) z so that backward sees
// exactly the values forward normalised.  Reductions are fixed-order two-stage (no float atomics).
#include <cstdlib>

#include "bf_common.cuh"

namespace {

constexpr int kFwdThreads = 256;
constexpr int kFwdWarps = kFwdThreads / 32;
constexpr int kFwdBlocksPerSm = 3;

template <typename T>
struct Pack8;
template <>
struct Pack8<__nv_bfloat16> {
    uint4 u;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = __ldcs(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void load_smem(const char* row, int c, int lane) {
        u = *reinterpret_cast<const uint4*>(row + c * 512 + lane * 16);
    }
    __device__ __forceinline__ void get(float (&v)[8]) const {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    // round to storage precision and keep the rounded values (what a later load would return)
    __device__ __forceinline__ void set(float (&v)[8]) {
        uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&b);
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ void store(__nv_bfloat16* p) const { __stcs(reinterpret_cast<uint4*>(p), u); }
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
    // fp32 is the parity mode: the 32-byte lane stride costs a 2-way bank conflict, accepted
    __device__ __forceinline__ void load_smem(const char* row, int c, int lane) {
        a = *reinterpret_cast<const float4*>(row + c * 1024 + lane * 32);
        b = *reinterpret_cast<const float4*>(row + c * 1024 + lane * 32 + 16);
    }
    __device__ __forceinline__ void get(float (&v)[8]) const {
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    }
    __device__ __forceinline__ void set(float (&v)[8]) {
        a = make_float4(v[0], v[1], v[2], v[3]);
        b = make_float4(v[4], v[5], v[6], v[7]);
    }
    __device__ __forceinline__ void store(float* p) const {
        __stcs(reinterpret_cast<float4*>(p), a);
        __stcs(reinterpret_cast<float4*>(p) + 1, b);
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

struct DropSpec {
    uint32_t k0, k1;     // Philox key = seed
    uint32_t site;       // counter word 2: 0x80000000 | site id (eps streams use tensor ids < 2^31 there)
    uint32_t step;       // counter word 3 (+ the device step counter)
    uint32_t threshold;  // keep iff u16 >= threshold; 0 = dropout off
    float scale;         // 1 / (1 - p)
    const uint32_t* step_ptr;
    uint32_t* keep;      // optional [rows][32]: per (row, lane) one word with the keep bytes of the lane's C octets (byte c =
                         // octet c * 32 + lane), written by forward, read by backward (the backward is issue-bound;
                         // regenerating the mask is half of its instructions)
};

// keep-multipliers (0 or scale) of the 8 elements of octet `oct` (= flat element index / 8)
__device__ __forceinline__ void drop_mult8(const DropSpec& d, uint32_t step, uint64_t oct, float (&m)[8]) {
    const uint4 r = bf_philox4x32_10((uint32_t)oct, (uint32_t)(oct >> 32), d.site, step, d.k0, d.k1);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[2 * i] = (w[i] & 0xffffu) >= d.threshold ? d.scale : 0.0f;
        m[2 * i + 1] = (w[i] >> 16) >= d.threshold ? d.scale : 0.0f;
    }
}

// multipliers and the keep byte (bit j = element j kept) from the same compares
__device__ __forceinline__ uint32_t drop_mult8_byte(const DropSpec& d, uint32_t step, uint64_t oct, float (&m)[8]) {
    const uint4 r = bf_philox4x32_10((uint32_t)oct, (uint32_t)(oct >> 32), d.site, step, d.k0, d.k1);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool k0 = (w[i] & 0xffffu) >= d.threshold, k1 = (w[i] >> 16) >= d.threshold;
        m[2 * i] = k0 ? d.scale : 0.0f;
        m[2 * i + 1] = k1 ? d.scale : 0.0f;
        if (k0) bits |= 1u << (2 * i);
        if (k1) bits |= 2u << (2 * i);
    }
    return bits;
}
// the same from the stored keep byte of the octet (bit j = element j kept)
__device__ __forceinline__ void drop_mult8_bits(const DropSpec& d, uint32_t bits, float (&m)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = (bits & (1u << j)) ? d.scale : 0.0f;
}
// ------------------------------------------------------------------ forward
template <typename T, int C>
__global__ void __launch_bounds__(kFwdThreads, kFwdBlocksPerSm)
    resln_fwd_kernel(const T* __restrict__ h, const T* __restrict__ r, const float* __restrict__ gamma,
                     const float* __restrict__ beta, T* __restrict__ z, T* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int64_t M, int64_t affine_stride, float eps, DropSpec drop) {
    constexpr int H = 256 * C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.y;
    const float* g = gamma + (int64_t)s * affine_stride;
    const float* b = beta ? beta + (int64_t)s * affine_stride : nullptr;
    const int64_t row0 = (int64_t)s * M;
    const uint32_t step = drop.step + (drop.step_ptr ? *drop.step_ptr : 0u);
    for (int64_t m = (int64_t)blockIdx.x * kFwdWarps + warp; m < M; m += (int64_t)gridDim.x * kFwdWarps) {
        const int64_t row = row0 + m;
        Pack8<T> pz[C];
        uint32_t kw = 0;  // keep bytes of this lane's C octets
        {
            Pack8<T> ph[C], pr[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                ph[c].load(h + row * H + c * 256 + lane * 8);
                pr[c].load(r + row * H + c * 256 + lane * 8);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float hv[8], rv[8], zv[8];
                ph[c].get(hv);
                pr[c].get(rv);
                if (drop.threshold) {
                    float mk[8];
                    kw |= drop_mult8_byte(drop, step, (uint64_t)row * (H / 8) + c * 32 + lane, mk) << (8 * c);
#pragma unroll
                    for (int j = 0; j < 8; ++j) zv[j] = fmaf(hv[j], mk[j], rv[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) zv[j] = hv[j] + rv[j];
                }
                pz[c].set(zv);
                pz[c].store(z + row * H + c * 256 + lane * 8);
            }
        }
        if (drop.threshold && drop.keep) drop.keep[row * 32 + lane] = kw;  // one coalesced 128 B store per row
        float sum = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float v[8];
            pz[c].get(v);
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += v[j];
        }
        const float mean = bf_warp_sum(sum) * (1.0f / H);
        float sq = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float v[8];
            pz[c].get(v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = v[j] - mean;
                sq = fmaf(d, d, sq);
            }
        }
        const float rstd = rsqrtf(bf_warp_sum(sq) * (1.0f / H) + eps);
        if (lane == 0) {
            mean_out[row] = mean;
            rstd_out[row] = rstd;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float v[8], gv[8], o[8];
            pz[c].get(v);
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
            Pack8<T>::store(y + row * H + c * 256 + lane * 8, o);
        }
    }
}

// ------------------------------------------------------------------ backward
// Tail shared by both backward kernels: block reduction of the three column accumulators through shared
// memory (kRedRows warps' worth at a time, fixed order), per-block partial, then the last block of each sample
// (and, for a shared affine, the last of those) adds the partials in a fixed order.  `red_raw`: kRedRows*3*H floats.
template <int C>
struct RedCfg {
    static constexpr int kRedRows = C <= 3 ? 4 : 2;  // <= 48 KB
};

template <int C, int kThreads>
__device__ __forceinline__ void bwd_finish(float (&acc_g)[C][8], float (&acc_b)[C][8], float (&acc_h)[C][8],
                                           float* red_raw, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                           float* __restrict__ dbias, float* __restrict__ partial,
                                           float* __restrict__ sample_part, unsigned int* __restrict__ counters,
                                           int64_t affine_stride, int S) {
    constexpr int kRedRows = RedCfg<C>::kRedRows;
    // ---- block reduction through shared memory, 4 warps' worth at a time; fixed order -> deterministic
    constexpr int H = 256 * C;
    constexpr int kWarps = kThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.y, nblk = gridDim.x;
    float(*red)[3 * H] = reinterpret_cast<float(*)[3 * H]>(red_raw);
    __shared__ bool is_last;
#pragma unroll 1
    for (int base = kRedRows; base < kWarps + kRedRows; base += kRedRows) {
        // round `base`: writers are warps [base, base+4) -- or, in the last round, warps [1, 4) folding into warp 0
        const bool last = base >= kWarps;
        const int w_lo = last ? 1 : base, w_hi = last ? kRedRows : base + kRedRows;
        if (warp >= w_lo && warp < w_hi) {
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    red[warp - w_lo][c * 256 + lane * 8 + j] = acc_g[c][j];
                    red[warp - w_lo][H + c * 256 + lane * 8 + j] = acc_b[c][j];
                    red[warp - w_lo][2 * H + c * 256 + lane * 8 + j] = acc_h[c][j];
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
                        acc_h[c][j] += red[warp][2 * H + c * 256 + lane * 8 + j];
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
                        acc_h[c][j] += red[r][2 * H + c * 256 + lane * 8 + j];
                    }
        }
        __syncthreads();
    }
    float* my_part = partial + ((int64_t)s * nblk + blockIdx.x) * 3 * H;
    if (warp == 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float* p0 = my_part + c * 256 + lane * 8;
            *reinterpret_cast<float4*>(p0) = make_float4(acc_g[c][0], acc_g[c][1], acc_g[c][2], acc_g[c][3]);
            *reinterpret_cast<float4*>(p0 + 4) = make_float4(acc_g[c][4], acc_g[c][5], acc_g[c][6], acc_g[c][7]);
            *reinterpret_cast<float4*>(p0 + H) = make_float4(acc_b[c][0], acc_b[c][1], acc_b[c][2], acc_b[c][3]);
            *reinterpret_cast<float4*>(p0 + H + 4) = make_float4(acc_b[c][4], acc_b[c][5], acc_b[c][6], acc_b[c][7]);
            *reinterpret_cast<float4*>(p0 + 2 * H) = make_float4(acc_h[c][0], acc_h[c][1], acc_h[c][2], acc_h[c][3]);
            *reinterpret_cast<float4*>(p0 + 2 * H + 4) = make_float4(acc_h[c][4], acc_h[c][5], acc_h[c][6], acc_h[c][7]);
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
    // ---- per-sample fixed-order pass by the last block of sample s: 16-byte L2 loads (ld.global.cg: the partials
    // were written by other SMs), 8 in flight per thread
    const bool shared_affine = affine_stride == 0 && S > 1;
    constexpr int kCol4 = 3 * H / 4;
    const float4* base4 = reinterpret_cast<const float4*>(partial + (int64_t)s * nblk * 3 * H);
    for (int c4 = threadIdx.x; c4 < kCol4; c4 += kThreads) {
        float4 t = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll 8
        for (int k = 0; k < nblk; ++k) {
            const float4 v = __ldcg(base4 + (int64_t)k * kCol4 + c4);
            t.x += v.x, t.y += v.y, t.z += v.z, t.w += v.w;
        }
        const int col = c4 * 4;  // H % 4 == 0: a float4 never straddles the gamma / beta / bias thirds
        if (col >= 2 * H) {
            if (dbias) *reinterpret_cast<float4*>(dbias + (int64_t)s * H + col - 2 * H) = t;
        } else if (shared_affine) {
            *reinterpret_cast<float4*>(sample_part + (int64_t)s * 2 * H + col) = t;
        } else if (col < H) {
            *reinterpret_cast<float4*>(dgamma + (int64_t)s * H + col) = t;
        } else if (dbeta) {
            *reinterpret_cast<float4*>(dbeta + (int64_t)s * H + col - H) = t;
        }
    }
    if (threadIdx.x == 0) counters[s] = 0u;
    if (!shared_affine) return;
    // ---- shared affine: the last of the S per-sample finishers adds the S per-sample sums in order
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(counters + S, 1u);
        is_last = (done == (unsigned int)S - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const float4* sp4 = reinterpret_cast<const float4*>(sample_part);
    for (int c4 = threadIdx.x; c4 < 2 * H / 4; c4 += kThreads) {
        float4 t = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        for (int k = 0; k < S; ++k) {
            const float4 v = __ldcg(sp4 + (int64_t)k * (2 * H / 4) + c4);
            t.x += v.x, t.y += v.y, t.z += v.z, t.w += v.w;
        }
        const int col = c4 * 4;
        if (col < H) *reinterpret_cast<float4*>(dgamma + col) = t;
        else if (dbeta) *reinterpret_cast<float4*>(dbeta + col - H) = t;
    }
    if (threadIdx.x == 0) counters[S] = 0u;
}


// workspace: [S + 1 counters, padded to 256 B][S][nblk][3][H] block partials [S][2][H] per-sample affine sums
template <int C>
struct BwdCfg {
    static constexpr int kThreads = C <= 2 ? 384 : 256;  // register budget: 3 x C x 8 accumulators per thread
};

template <typename T, int C, bool kDrop>
__global__ void __launch_bounds__(BwdCfg<C>::kThreads, 1)
    resln_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ z, const float* __restrict__ gamma,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in, T* __restrict__ dz,
                     T* __restrict__ dh, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ dbias, float* __restrict__ partial, float* __restrict__ sample_part,
                     unsigned int* __restrict__ counters, int64_t M, int64_t affine_stride, int S, DropSpec drop) {
    constexpr int H = 256 * C;
    constexpr int kThreads = BwdCfg<C>::kThreads;
    constexpr int kWarps = kThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.y, nblk = gridDim.x;
    const float* g = gamma + (int64_t)s * affine_stride;
    const int64_t row0 = (int64_t)s * M;
    const uint32_t step = drop.step + (drop.step_ptr ? *drop.step_ptr : 0u);
    float acc_g[C][8], acc_b[C][8], acc_h[C][8];  // sum gy*xhat, sum gy, sum dh
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc_g[c][j] = acc_b[c][j] = acc_h[c][j] = 0.0f;

    const int64_t m_step = (int64_t)nblk * kWarps;
    int64_t m = (int64_t)blockIdx.x * kWarps + warp;
    Pack8<T> nz[C], ng[C];  // the next row's loads are in flight while this row is reduced
    if (m < M) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            nz[c].load(z + (row0 + m) * H + c * 256 + lane * 8);
            ng[c].load(gy + (row0 + m) * H + c * 256 + lane * 8);
        }
    }
    float mean_n = 0.0f, rstd_n = 0.0f;  // the row statistics travel one iteration ahead as well
    if (m < M) mean_n = __ldg(mean_in + row0 + m), rstd_n = __ldg(rstd_in + row0 + m);
    for (; m < M; m += m_step) {
        const int64_t row = row0 + m;
        Pack8<T> pz[C], pg[C];
#pragma unroll
        for (int c = 0; c < C; ++c) pz[c] = nz[c], pg[c] = ng[c];
        const float mean = mean_n, rstd = rstd_n;
        if (m + m_step < M) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                nz[c].load(z + (row + m_step) * H + c * 256 + lane * 8);
                ng[c].load(gy + (row + m_step) * H + c * 256 + lane * 8);
            }
            mean_n = __ldg(mean_in + row + m_step), rstd_n = __ldg(rstd_in + row + m_step);
        }
        float s1 = 0.0f, s2 = 0.0f;  // sum a, sum a*xhat with a = gy*gamma
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float zv[8], gv[8], gm[8];
            pz[c].get(zv);
            pg[c].get(gv);
            ld8f(g + c * 256 + lane * 8, gm);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xh = (zv[j] - mean) * rstd;
                const float a = gv[j] * gm[j];
                s1 += a;
                s2 = fmaf(a, xh, s2);
                acc_g[c][j] = fmaf(gv[j], xh, acc_g[c][j]);
                acc_b[c][j] += gv[j];
            }
        }
        const float c1 = bf_warp_sum(s1) * (1.0f / H), c2 = bf_warp_sum(s2) * (1.0f / H);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float zv[8], gv[8], gm[8], o[8];
            pz[c].get(zv);
            pg[c].get(gv);
            ld8f(g + c * 256 + lane * 8, gm);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xh = (zv[j] - mean) * rstd;
                o[j] = rstd * (gv[j] * gm[j] - c1 - xh * c2);
            }
            Pack8<T>::store(dz + row * H + c * 256 + lane * 8, o);
            if (kDrop) {
                float mk[8];
                if (drop.keep) drop_mult8_bits(drop, (__ldg(drop.keep + row * 32 + lane) >> (8 * c)) & 0xffu, mk);
                else drop_mult8(drop, step, (uint64_t)row * (H / 8) + c * 32 + lane, mk);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] *= mk[j];
                Pack8<T>::store(dh + row * H + c * 256 + lane * 8, o);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc_h[c][j] += o[j];
        }
    }

    constexpr int kRedRows = RedCfg<C>::kRedRows;
    __shared__ float red[kRedRows * 3 * H];
    bwd_finish<C, kThreads>(acc_g, acc_b, acc_h, red, dgamma, dbeta, dbias, partial, sample_part, counters, affine_stride,
                            S);
}

// ---- staged backward: the rows travel global -> shared memory as 1-D bulk async copies (TMA unit, completion on an
// mbarrier) into a per-warp ring of kStages rows, so the bytes in flight no longer depend on registers
// (12 warps x kStages x 2 rows of H elements, ~100+ KB per SM) and 12 warps fit the register file
template <typename T, int C>
struct StagedCfg {
    // C >= 3: 72+ accumulators, two packed rows and the prefetched row statistics need > 168 registers, i.e. at most
    // 2 warps per SM sub-partition (with 12 warps ptxas spilled the prefetched mean / rstd right behind their loads,
    // which turned the prefetch back into a stall: ncu source page, 16 % of the samples on that STL)
    static constexpr int kThreads = C <= 2 ? 384 : 256;
    static constexpr int kWarps = kThreads / 32;
    static constexpr int kRowBytes = 256 * C * (int)sizeof(T);
    static constexpr int kStages = sizeof(T) == 2 ? 4 : 2;
    static constexpr int kRingBytes = kWarps * kStages * 2 * kRowBytes;
    static constexpr int kRedBytes = RedCfg<C>::kRedRows * 3 * 256 * C * (int)sizeof(float);
    static constexpr int kBarBytes = kWarps * kStages * 8;
    static constexpr int kSmemBytes = (kRingBytes > kRedBytes ? kRingBytes : kRedBytes) + kBarBytes;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rl_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void rl_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rl_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void rl_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <typename T, int C, bool kDrop>
__global__ void __launch_bounds__(StagedCfg<T, C>::kThreads, 1)
    resln_bwd_staged_kernel(const T* __restrict__ gy, const T* __restrict__ z, const float* __restrict__ gamma,
                            const float* __restrict__ mean_in, const float* __restrict__ rstd_in, T* __restrict__ dz,
                            T* __restrict__ dh, float* __restrict__ dgamma, float* __restrict__ dbeta,
                            float* __restrict__ dbias, float* __restrict__ partial, float* __restrict__ sample_part,
                            unsigned int* __restrict__ counters, int64_t M, int64_t affine_stride, int S, DropSpec drop) {
    using Cfg = StagedCfg<T, C>;
    constexpr int H = 256 * C;
    constexpr int kWarps = Cfg::kWarps, kStages = Cfg::kStages, kRowBytes = Cfg::kRowBytes;
    extern __shared__ __align__(128) char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.y, nblk = gridDim.x;
    const float* g = gamma + (int64_t)s * affine_stride;
    const int64_t row0 = (int64_t)s * M;
    const uint32_t step = drop.step + (drop.step_ptr ? *drop.step_ptr : 0u);
    char* ring = smem + (size_t)warp * kStages * 2 * kRowBytes;  // this warp's [kStages][z row, gy row]
    const uint32_t bars = smem_u32(smem + (Cfg::kSmemBytes - Cfg::kBarBytes)) + warp * kStages * 8;
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < kStages; ++st) rl_mbar_init(bars + st * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // column accumulators as packed fp32 pairs (FFMA2 / FADD2: two IEEE fp32 operations per issued instruction --
    // this kernel is issue-bound, not HBM-bound, see profiles/): sum gy*xhat, sum gy, sum dh
    bf_f2 acc_g2[C][4], acc_b2[C][4], acc_h2[C][4];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc_g2[c][j] = acc_b2[c][j] = acc_h2[c][j] = bf_splat2(0.0f);

    const int64_t m_step = (int64_t)nblk * kWarps;
    const int64_t m_first = (int64_t)blockIdx.x * kWarps + warp;
    if (lane == 0) {  // prologue: fill the ring
#pragma unroll
        for (int st = 0; st < kStages; ++st) {
            const int64_t mm = m_first + st * m_step;
            if (mm < M) {
                rl_mbar_expect_tx(bars + st * 8, 2 * kRowBytes);
                rl_bulk_load(smem_u32(ring + (st * 2) * kRowBytes), z + (row0 + mm) * H, kRowBytes, bars + st * 8);
                rl_bulk_load(smem_u32(ring + (st * 2 + 1) * kRowBytes), gy + (row0 + mm) * H, kRowBytes, bars + st * 8);
            }
        }
    }
    int st = 0;
    uint32_t parity = 0;
    float mean_n = 0.0f, rstd_n = 0.0f;  // row statistics: plain loads issued one iteration ahead
    uint32_t kb_n = 0;                   // ... and, when forward stored it, the word with the keep bytes of this lane's octets
    const bool stored_mask = kDrop && drop.keep != nullptr;
    auto fetch_keep = [&](int64_t r) { kb_n = __ldg(drop.keep + r * 32 + lane); };
    if (m_first < M) {
        mean_n = __ldg(mean_in + row0 + m_first), rstd_n = __ldg(rstd_in + row0 + m_first);
        if (stored_mask) fetch_keep(row0 + m_first);
    }
    for (int64_t m = m_first; m < M; m += m_step) {
        const int64_t row = row0 + m;
        const float mean = mean_n, rstd = rstd_n;
        const uint32_t kb = kb_n;
        if (m + m_step < M) {
            mean_n = __ldg(mean_in + row + m_step), rstd_n = __ldg(rstd_in + row + m_step);
            if (stored_mask) fetch_keep(row + m_step);
        }
        rl_mbar_wait(bars + st * 8, parity);
        const char* zr = ring + (st * 2) * kRowBytes;
        const char* gr = zr + kRowBytes;
        Pack8<T> pz[C], pg[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            pz[c].load_smem(zr, c, lane);
            pg[c].load_smem(gr, c, lane);
        }
        __syncwarp();  // every lane has its copy of the row in registers: the slot can be refilled
        if (lane == 0) {
            const int64_t mn = m + (int64_t)kStages * m_step;
            if (mn < M) {
                rl_mbar_expect_tx(bars + st * 8, 2 * kRowBytes);
                rl_bulk_load(smem_u32(ring + (st * 2) * kRowBytes), z + (row0 + mn) * H, kRowBytes, bars + st * 8);
                rl_bulk_load(smem_u32(ring + (st * 2 + 1) * kRowBytes), gy + (row0 + mn) * H, kRowBytes, bars + st * 8);
            }
        }
        if (++st == kStages) st = 0, parity ^= 1u;

        // pass 1: xhat and a = gy*gamma stay in registers (packed), row sums of a and a*xhat
        const bf_f2 nmean2 = bf_splat2(-mean), rstd2 = bf_splat2(rstd);
        bf_f2 xh2[C][4], a2[C][4];
        bf_f2 s1_2 = bf_splat2(0.0f), s2_2 = bf_splat2(0.0f);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float zv[8], gv[8], gm[8];
            pz[c].get(zv);
            pg[c].get(gv);
            ld8f(g + c * 256 + lane * 8, gm);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bf_f2 g2 = bf_pack2(gv[2 * j], gv[2 * j + 1]);
                xh2[c][j] = bf_mul2(bf_add2(bf_pack2(zv[2 * j], zv[2 * j + 1]), nmean2), rstd2);  // (z - mean) * rstd
                a2[c][j] = bf_mul2(g2, bf_pack2(gm[2 * j], gm[2 * j + 1]));
                s1_2 = bf_add2(s1_2, a2[c][j]);
                s2_2 = bf_fma2(a2[c][j], xh2[c][j], s2_2);
                acc_g2[c][j] = bf_fma2(g2, xh2[c][j], acc_g2[c][j]);
                acc_b2[c][j] = bf_add2(acc_b2[c][j], g2);
            }
        }
        float s1a, s1b, s2a, s2b;
        bf_unpack2(s1_2, s1a, s1b);
        bf_unpack2(s2_2, s2a, s2b);
        const float c1 = bf_warp_sum(s1a + s1b) * (1.0f / H), c2 = bf_warp_sum(s2a + s2b) * (1.0f / H);
        // pass 2: dz = rstd * (a - c1 - xhat*c2), dh = dz * keep / (1 - p)
        const bf_f2 nc1_2 = bf_splat2(-c1), nc2_2 = bf_splat2(-c2);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float o[8];
            bf_f2 o2[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o2[j] = bf_mul2(bf_fma2(xh2[c][j], nc2_2, bf_add2(a2[c][j], nc1_2)), rstd2);
                bf_unpack2(o2[j], o[2 * j], o[2 * j + 1]);
            }
            Pack8<T>::store(dz + row * H + c * 256 + lane * 8, o);
            if (kDrop) {
                float mk[8];
                if (stored_mask) drop_mult8_bits(drop, (kb >> (8 * c)) & 0xffu, mk);
                else drop_mult8(drop, step, (uint64_t)row * (H / 8) + c * 32 + lane, mk);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    o2[j] = bf_mul2(o2[j], bf_pack2(mk[2 * j], mk[2 * j + 1]));
                    bf_unpack2(o2[j], o[2 * j], o[2 * j + 1]);
                }
                Pack8<T>::store(dh + row * H + c * 256 + lane * 8, o);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) acc_h2[c][j] = bf_add2(acc_h2[c][j], o2[j]);
        }
    }
    float acc_g[C][8], acc_b[C][8], acc_h[C][8];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bf_unpack2(acc_g2[c][j], acc_g[c][2 * j], acc_g[c][2 * j + 1]);
            bf_unpack2(acc_b2[c][j], acc_b[c][2 * j], acc_b[c][2 * j + 1]);
            bf_unpack2(acc_h2[c][j], acc_h[c][2 * j], acc_h[c][2 * j + 1]);
        }
    __syncthreads();  // all warps are done with their rings: the reduction scratch aliases them
    bwd_finish<C, Cfg::kThreads>(acc_g, acc_b, acc_h, reinterpret_cast<float*>(smem), dgamma, dbeta, dbias, partial,
                                 sample_part, counters, affine_stride, S);
}

inline int bwd_blocks_w(int64_t S, int64_t M, int kWarps) {
    int64_t per = bf_num_sms() / S;  // one block per SM, spread over the samples
    if (per < 1) per = 1;
    const int64_t need = (M + kWarps - 1) / kWarps;
    if (per > need) per = need;
    return (int)(per < 1 ? 1 : per);
}
template <int C>
inline int bwd_blocks(int64_t S, int64_t M) {
    return bwd_blocks_w(S, M, BwdCfg<C>::kThreads / 32);
}
// upper bound over both backward kernels (the workspace is sized with it): fewer warps per block -> more blocks
inline int bwd_blocks_h(int64_t S, int64_t M, int64_t H) { return bwd_blocks_w(S, M, H / 256 <= 2 ? 12 : 8); }
static_assert(StagedCfg<float, 4>::kSmemBytes <= 227 * 1024, "staged ring exceeds shared memory");
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t counters_bytes(int64_t S) { return (((S + 1) * 4 + 255) / 256) * 256; }

DropSpec make_drop(float p, uint64_t seed, uint32_t step, uint32_t site) {
    DropSpec d;
    d.k0 = (uint32_t)(seed & 0xffffffffu), d.k1 = (uint32_t)(seed >> 32);
    d.site = 0x80000000u | site;
    d.step = step;
    d.step_ptr = bf_step_counter();
    double t = (double)p * 65536.0 + 0.5;
    d.threshold = p <= 0.0f ? 0u : (uint32_t)(t > 65535.0 ? 65535.0 : t);
    d.scale = p <= 0.0f ? 1.0f : 1.0f / (1.0f - p);
    d.keep = nullptr;
    return d;
}

template <typename T, int C>
int launch_fwd_c(const void* h, const void* r, const float* gamma, const float* beta, void* z, void* y, float* mean,
                 float* rstd, int64_t S, int64_t M, int64_t astride, float eps, const DropSpec& d, cudaStream_t st) {
    int64_t need = (M + kFwdWarps - 1) / kFwdWarps;
    int64_t cap = (int64_t)bf_num_sms() * kFwdBlocksPerSm / S;  // one full wave of resident blocks over all samples
    if (cap < 1) cap = 1;
    if (need > cap) need = cap;
    dim3 grid((unsigned)(need < 1 ? 1 : need), (unsigned)S);
    resln_fwd_kernel<T, C><<<grid, kFwdThreads, 0, st>>>(reinterpret_cast<const T*>(h), reinterpret_cast<const T*>(r),
                                                         gamma, beta, reinterpret_cast<T*>(z), reinterpret_cast<T*>(y),
                                                         mean, rstd, M, astride, eps, d);
    return 0;
}

inline bool use_staged_bwd() {
    // bf_set_option(BF_OPT_RESLN_BWD_STAGED, 0): force the register-prefetch kernel (A/B timing, debugging)
    return bf_option(BF_OPT_RESLN_BWD_STAGED) != 0;
}

template <typename T, int C>
int launch_bwd_c(const void* gy, const void* z, const float* gamma, const float* mean, const float* rstd, void* dz,
                 void* dh, float* dgamma, float* dbeta, float* dbias, void* ws, int64_t S, int64_t M, int64_t astride,
                 const DropSpec& d, cudaStream_t st) {
    constexpr int H = 256 * C;
    using SC = StagedCfg<T, C>;
    const bool staged = use_staged_bwd();  // BF_OPT_RESLN_BWD_STAGED = 0 keeps the register-prefetch kernel selectable (A/B timing)
    const int nblk = staged ? bwd_blocks_w(S, M, SC::kWarps) : bwd_blocks<C>(S, M);
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
    float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + counters_bytes(S));
    float* sample_part = partial + S * (int64_t)nblk * 3 * H;
    dim3 grid((unsigned)nblk, (unsigned)S);
    const T* gyp = reinterpret_cast<const T*>(gy);
    const T* zp = reinterpret_cast<const T*>(z);
    T* dzp = reinterpret_cast<T*>(dz);
    T* dhp = reinterpret_cast<T*>(dh);
    if (staged) {
        static bool attr_done[2] = {false, false};
        auto* k1 = resln_bwd_staged_kernel<T, C, true>;
        auto* k0 = resln_bwd_staged_kernel<T, C, false>;
        const int which = d.threshold ? 1 : 0;
        if (!attr_done[which]) {
            BF_CUDA_OK(cudaFuncSetAttribute(which ? k1 : k0, cudaFuncAttributeMaxDynamicSharedMemorySize, SC::kSmemBytes));
            attr_done[which] = true;
        }
        if (d.threshold)
            k1<<<grid, SC::kThreads, SC::kSmemBytes, st>>>(gyp, zp, gamma, mean, rstd, dzp, dhp, dgamma, dbeta, dbias, partial,
                                                          sample_part, counters, M, astride, (int)S, d);
        else
            k0<<<grid, SC::kThreads, SC::kSmemBytes, st>>>(gyp, zp, gamma, mean, rstd, dzp, nullptr, dgamma, dbeta, dbias,
                                                          partial, sample_part, counters, M, astride, (int)S, d);
        return 0;
    }
    if (d.threshold)
        resln_bwd_kernel<T, C, true><<<grid, BwdCfg<C>::kThreads, 0, st>>>(gyp, zp, gamma, mean, rstd, dzp, dhp, dgamma,
                                                                           dbeta, dbias, partial, sample_part, counters, M,
                                                                           astride, (int)S, d);
    else
        resln_bwd_kernel<T, C, false><<<grid, BwdCfg<C>::kThreads, 0, st>>>(gyp, zp, gamma, mean, rstd, dzp, nullptr, dgamma,
                                                                            dbeta, dbias, partial, sample_part, counters, M,
                                                                            astride, (int)S, d);
    return 0;
}

#define BF_RESLN_CASE(FN, T, CC, ...)               \
    case CC: {                                      \
        const int rc_ = FN<T, CC>(__VA_ARGS__);     \
        if (rc_ != 0) return rc_;                   \
    } break;
#define BF_RESLN_DISPATCH(FN, T, ...)                 \
    switch (H / 256) {                                \
        BF_RESLN_CASE(FN, T, 1, __VA_ARGS__)          \
        BF_RESLN_CASE(FN, T, 2, __VA_ARGS__)          \
        BF_RESLN_CASE(FN, T, 3, __VA_ARGS__)          \
        BF_RESLN_CASE(FN, T, 4, __VA_ARGS__)          \
        default: bf_set_error("resln: H/256 must be 1..4"); return BF_ERR_UNSUPPORTED; \
    }

__global__ void dropout_mask_kernel(uint8_t* __restrict__ out, int64_t n, DropSpec drop) {
    const uint32_t step = drop.step + (drop.step_ptr ? *drop.step_ptr : 0u);
    const int64_t n_oct = (n + 7) >> 3;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_oct; o += (int64_t)gridDim.x * blockDim.x) {
        float mk[8];
        if (drop.threshold) {
            drop_mult8(drop, step, (uint64_t)o, mk);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) mk[j] = 1.0f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (o * 8 + j < n) out[o * 8 + j] = mk[j] != 0.0f ? 1 : 0;
    }
}

}  // namespace

extern "C" int bf_resln_supported(int64_t H) { return (H % 256 == 0 && H >= 256 && H <= 1024) ? 1 : 0; }

extern "C" int bf_resln_fwd_keep(const void* h, const void* r, int32_t dtype, const float* gamma, const float* beta,
                                 int64_t affine_stride, int64_t S, int64_t M, int64_t H, float eps, float p_drop,
                                 uint64_t seed, uint32_t step, uint32_t site_id, void* z, void* y, float* mean, float* rstd,
                                 uint32_t* keep, void* stream);
extern "C" int bf_resln_fwd(const void* h, const void* r, int32_t dtype, const float* gamma, const float* beta,
                            int64_t affine_stride, int64_t S, int64_t M, int64_t H, float eps, float p_drop,
                            uint64_t seed, uint32_t step, uint32_t site_id, void* z, void* y, float* mean, float* rstd,
                            void* stream) {
    return bf_resln_fwd_keep(h, r, dtype, gamma, beta, affine_stride, S, M, H, eps, p_drop, seed, step, site_id, z, y, mean,
                             rstd, nullptr, stream);
}
extern "C" int bf_resln_fwd_keep(const void* h, const void* r, int32_t dtype, const float* gamma, const float* beta,
                                 int64_t affine_stride, int64_t S, int64_t M, int64_t H, float eps, float p_drop,
                                 uint64_t seed, uint32_t step, uint32_t site_id, void* z, void* y, float* mean, float* rstd,
                                 uint32_t* keep, void* stream) {
    BF_CHECK_ARG(h && r && gamma && z && y && mean && rstd, "null pointer");
    BF_CHECK_ARG(dtype == BF_F32 || dtype == BF_BF16, "bad dtype");
    BF_CHECK_ARG(S >= 1 && M >= 0 && S <= 65535, "bad S or M");
    BF_CHECK_ARG(bf_resln_supported(H), "H must be 256, 512, 768 or 1024");
    BF_CHECK_ARG(affine_stride == 0 || affine_stride >= H, "bad affine stride");
    BF_CHECK_ARG(p_drop >= 0.0f && p_drop < 1.0f, "p_drop must be in [0, 1)");
    BF_CHECK_ARG(site_id < 0x80000000u, "site_id must be < 2^31");
    BF_CHECK_ARG(aligned16(h) && aligned16(r) && aligned16(z) && aligned16(y) && aligned16(gamma) && (!beta || aligned16(beta)),
                 "h, r, z, y, gamma, beta must be 16-byte aligned (vector loads / stores)");
    if (M == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DropSpec d = make_drop(p_drop, seed, step, site_id);
    d.keep = keep;
    if (dtype == BF_BF16) {
        BF_RESLN_DISPATCH(launch_fwd_c, __nv_bfloat16, h, r, gamma, beta, z, y, mean, rstd, S, M, affine_stride, eps, d, st);
    } else {
        BF_RESLN_DISPATCH(launch_fwd_c, float, h, r, gamma, beta, z, y, mean, rstd, S, M, affine_stride, eps, d, st);
    }
    BF_LAUNCH_OK();
    return 0;
}

extern "C" int64_t bf_resln_bwd_workspace_bytes(int64_t S, int64_t M, int64_t H) {
    if (S < 1) S = 1;
    const int64_t nblk = bwd_blocks_h(S, M < 1 ? 1 : M, H);
    return counters_bytes(S) + (S * nblk * 3 * H + S * 2 * H) * (int64_t)sizeof(float);
}

extern "C" int bf_resln_bwd_keep(const void* gy, const void* z, int32_t dtype, const float* gamma, int64_t affine_stride,
                                 const float* mean, const float* rstd, int64_t S, int64_t M, int64_t H, float p_drop,
                                 uint64_t seed, uint32_t step, uint32_t site_id, void* dz, void* dh, float* dgamma,
                                 float* dbeta, float* dbias, void* workspace, const uint32_t* keep, void* stream);
extern "C" int bf_resln_bwd(const void* gy, const void* z, int32_t dtype, const float* gamma, int64_t affine_stride,
                            const float* mean, const float* rstd, int64_t S, int64_t M, int64_t H, float p_drop,
                            uint64_t seed, uint32_t step, uint32_t site_id, void* dz, void* dh, float* dgamma,
                            float* dbeta, float* dbias, void* workspace, void* stream) {
    return bf_resln_bwd_keep(gy, z, dtype, gamma, affine_stride, mean, rstd, S, M, H, p_drop, seed, step, site_id, dz, dh, dgamma,
                             dbeta, dbias, workspace, nullptr, stream);
}
extern "C" int bf_resln_bwd_keep(const void* gy, const void* z, int32_t dtype, const float* gamma, int64_t affine_stride,
                                 const float* mean, const float* rstd, int64_t S, int64_t M, int64_t H, float p_drop,
                                 uint64_t seed, uint32_t step, uint32_t site_id, void* dz, void* dh, float* dgamma,
                                 float* dbeta, float* dbias, void* workspace, const uint32_t* keep, void* stream) {
    BF_CHECK_ARG(gy && z && gamma && mean && rstd && dz && dgamma && workspace, "null pointer");
    BF_CHECK_ARG(dtype == BF_F32 || dtype == BF_BF16, "bad dtype");
    BF_CHECK_ARG(S >= 1 && M >= 1 && S <= 65535, "bad S or M");
    BF_CHECK_ARG(bf_resln_supported(H), "H must be 256, 512, 768 or 1024");
    BF_CHECK_ARG(affine_stride == 0 || affine_stride >= H, "bad affine stride");
    BF_CHECK_ARG(p_drop >= 0.0f && p_drop < 1.0f, "p_drop must be in [0, 1)");
    BF_CHECK_ARG(p_drop <= 0.0f || dh, "dh is required when p_drop > 0");
    BF_CHECK_ARG(site_id < 0x80000000u, "site_id must be < 2^31");
    BF_CHECK_ARG(aligned16(gy) && aligned16(z) && aligned16(dz) && (!dh || aligned16(dh)) && aligned16(gamma) &&
                     aligned16(dgamma) && (!dbeta || aligned16(dbeta)) && (!dbias || aligned16(dbias)) && aligned16(workspace),
                 "gy, z, dz, dh, gamma, dgamma, dbeta, dbias, workspace must be 16-byte aligned (vector / bulk copies)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DropSpec d = make_drop(p_drop, seed, step, site_id);
    d.keep = const_cast<uint32_t*>(keep);
    if (dtype == BF_BF16) {
        BF_RESLN_DISPATCH(launch_bwd_c, __nv_bfloat16, gy, z, gamma, mean, rstd, dz, dh, dgamma, dbeta, dbias, workspace, S,
                          M, affine_stride, d, st);
    } else {
        BF_RESLN_DISPATCH(launch_bwd_c, float, gy, z, gamma, mean, rstd, dz, dh, dgamma, dbeta, dbias, workspace, S, M,
                          affine_stride, d, st);
    }
    BF_LAUNCH_OK();
    return 0;
}

extern "C" int bf_dropout_mask(uint8_t* out, int64_t n, float p_drop, uint64_t seed, uint32_t step, uint32_t site_id,
                               void* stream) {
    BF_CHECK_ARG(out || n == 0, "null pointer");
    BF_CHECK_ARG(p_drop >= 0.0f && p_drop < 1.0f, "p_drop must be in [0, 1)");
    BF_CHECK_ARG(site_id < 0x80000000u, "site_id must be < 2^31");
    if (n <= 0) return 0;
    const DropSpec d = make_drop(p_drop, seed, step, site_id);
    const int64_t n_oct = (n + 7) >> 3;
    int64_t blocks = (n_oct + 255) / 256;
    const int64_t cap = (int64_t)bf_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    dropout_mask_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, n, d);
    BF_LAUNCH_OK();
    return 0;
}
