// Short-sequence self-attention of the HOST model around the Bayesian q / k / v / output projections
// (opt-in, `accelerate_host_(attention=True)`): softmax(q k^T * scale) -> dropout -> . v for whole sequences of
// T <= 128 tokens and head width 64, forward and backward, one (sequence, head) per thread block.
//
// Why it exists: with every Linear on the tensor cores, the library attention the host model calls
// (F.scaled_dot_product_attention) was the largest non-contraction block of the training step -- at T = 128 its fused
// kernels run at ~200 TFLOP/s and need three helper passes per layer (dO.O, dq conversion, dropout masks): 2.2 ms per
// BERT-base layer against 0.75 ms of HBM time for q, k, v, o and their gradients.  A sequence of 128 tokens fits one
// thread block, so there is no online-softmax recurrence, no split of the key axis and no second pass for dq:
//
//   forward : S = q k^T (registers) -> row softmax -> Philox keep mask -> O = P_d v;  O and the base-2 log-sum-exp kept
//   backward: S, P recomputed; dP = dO v^T; dS = P o (dP_d - D) * scale with D = sum_k P_d dP_d (== rowsum(dO o O), so O
//             is not read); P_d and dS staged in shared memory, then dv = P_d^T dO, dk = dS^T q and dq = dS k by the same
//             block -- every output is written exactly once, nothing is accumulated with atomics: deterministic.
//
// q, k, v are read in place from the projection outputs ([B, T, heads*64] rows, any strides with a unit inner stride),
// O and the gradients are written in the [B, T, heads, 64] layout the surrounding reshape expects, so no transposed
// copies are made.
//
// Two kernel families share this entry point and one keep mask.  T == 128 (BERT's GLUE configuration) runs on tcgen05
// (bf_attention_tc.cu: TMA-loaded tiles, TMEM accumulators, one row per thread: 0.40 + 0.60 ms per BERT-base layer).
// THIS file holds the kernels for the other lengths (16 <= T < 128) and the fallback (BF_OPT_ATTN_TC = 0): matrix
// products with mma.sync m16n8k16 (bf16 in, fp32 accumulate) on fragments loaded with ldmatrix, tiles double buffered
// with cp.async, a 16-warp backward (8 row groups x 2 key halves).  At T = 128 they take 0.67 + 1.60 ms: issue-bound
// (a third of the instructions are Philox, and every two MMAs need an ldmatrix).
//
// The dropout keep mask is a pure function of its counter (these kernels regenerate it in backward; the tcgen05
// forward also hands its keep bits to the tcgen05 backward).  keep(b, h, q, k) = u16 >= round(p * 65536) with the u16 taken from
// Philox4x32-10(counter = (row, (k % 8) / 2 + 4 * (k / 32), 0x40000000 | site, step [+ device step counter]),
// key = seed), row = (b * heads + h) * T + q: word (k / 8) % 4 of the result, low half for even k, high half for odd k
// -- i.e. one call yields the 8 values one thread holds of a query row in the MMA accumulator layout.
// bf_attention_dropout_mask writes the mask as bytes (tests).
#include "bf_common.cuh"

namespace attn {

constexpr int TMAX = 128, D = 64, kThreads = 256;  // 8 warps x 16 query rows (phase B of the backward: x 16 keys)
constexpr int TILE_BYTES = TMAX * D * 2;      // one [128, 64] bf16 tile: 16 KiB
constexpr int SQ_BYTES = TMAX * TMAX * 2;     // one [128, 128] bf16 tile: 32 KiB
constexpr int kBwdThreads = 512;              // backward: 16 warps = 8 row groups x 2 key halves
constexpr int FWD_BUF = 3 * TILE_BYTES;       // q, k, v of one pair
constexpr int FWD_SMEM = 2 * FWD_BUF;         // double buffered: 96 KiB, two blocks per SM
constexpr int BWD_BUF = 4 * TILE_BYTES;       // q, k, v, dO of one pair
constexpr int BWD_SMEM = 2 * BWD_BUF + 2 * SQ_BYTES + 2 * TMAX * 4;  // + P_d, dS, partial D: 193 KiB

struct Params {
    const __nv_bfloat16 *q, *k, *v, *o, *dout;
    __nv_bfloat16 *out, *dq, *dk, *dv;
    float* lse;  // [B, heads, T] base-2 log-sum-exp of the scaled scores
    int64_t q_sb, q_sh, q_st, k_sb, k_sh, k_st, v_sb, v_sh, v_st;  // element strides of batch, head, token
    int B, H, T;
    float scale, scale_log2e, inv_keep;
    uint32_t thresh;  // keep iff u16 >= thresh (0: dropout off)
    uint32_t k0, k1, step, site;
    const uint32_t* step_ptr;
};

// ------------------------------------------------------------------ small PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// d (16x8 fp32) += a (16x16 bf16, row) * b (16x8 bf16, col)
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                 "{%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// byte offset of element (row, col) in a [rows][64] bf16 tile (128 B rows, 16 B chunks XOR-swizzled by row % 8)
__device__ __forceinline__ uint32_t off64(int row, int col) {
    return (uint32_t)(row * 128 + ((((col >> 3) ^ row) & 7) << 4) + ((col & 7) << 1));
}
// same for a [rows][128] bf16 tile (256 B rows: the low 3 bits of the chunk index are swizzled)
__device__ __forceinline__ uint32_t off128(int row, int col) {
    const int ch = col >> 3;
    return (uint32_t)(row * 256 + (((ch & 8) | ((ch ^ row) & 7)) << 4) + ((col & 7) << 1));
}

// [T, 64] rows of one (batch, head) from global memory into a swizzled tile; rows >= T are zero-filled
__device__ __forceinline__ void load_tile(uint8_t* tile, const __nv_bfloat16* base, int64_t st, int T) {
    for (int i = threadIdx.x; i < TMAX * 8; i += blockDim.x) {
        const int r = i >> 3, ch = i & 7;
        uint8_t* dst = tile + r * 128 + (((ch ^ r) & 7) << 4);
        if (r < T) cp_async16(smem_u32(dst), base + (int64_t)r * st + ch * 8);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// keep bits of the values a thread holds of query row `row` in key tiles 4 * m0 .. 4 * (m0 + NM) - 1
// (local tile jl, element e -> bit 2 * jl + e); one Philox call covers 4 tiles
template <int NM>
__device__ __forceinline__ uint32_t keep_bits(const Params& p, uint32_t row, int c, uint32_t step, int m0) {
    uint32_t bits = 0;
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        const uint4 r = bf_philox4x32_10(row, (uint32_t)(c + 4 * (m0 + m)), 0x40000000u | p.site, step, p.k0, p.k1);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint32_t lo = w[t] & 0xffffu, hi = w[t] >> 16;
            bits |= (uint32_t)(lo >= p.thresh) << (2 * (4 * m + t));
            bits |= (uint32_t)(hi >= p.thresh) << (2 * (4 * m + t) + 1);
        }
    }
    return bits;
}

// S = A_w B^T for the 16 rows of this warp and 16 * NJP rows of B starting at key0: s[j] = accumulator of the
// 8 columns key0 + 8j .. key0 + 8j + 7.  A, B are [rows][64] tiles.
template <int NJP>
__device__ __forceinline__ void scores(float (&s)[2 * NJP][4], const uint8_t* sA, const uint8_t* sB, int row0, int key0,
                                       int lane) {
#pragma unroll
    for (int j = 0; j < 2 * NJP; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        ldsm_x4(a, a_base + off64(row0 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 16 + (lane >> 4) * 8));
#pragma unroll
        for (int jp = 0; jp < NJP; ++jp) {
            uint32_t b[4];
            ldsm_x4(b, b_base + off64(key0 + jp * 16 + (lane & 7) + (lane >> 4) * 8, kk * 16 + ((lane >> 3) & 1) * 8));
            mma16816(s[2 * jp], a, b[0], b[1]);
            mma16816(s[2 * jp + 1], a, b[2], b[3]);
        }
    }
}

// the 16 x 64 accumulator tile of a warp (o[j]: columns 8j + 2c, +1 of rows g and g + 8) -> its own 16 rows of a
// [rows][64] staging tile -> 128 B rows of global memory, 16 B per lane
__device__ __forceinline__ void store_rows64(uint8_t* stage, int row0, const float (&o)[8][4], float r0, float r1,
                                             __nv_bfloat16* gbase, int64_t g_st, int lane) {
    const int g = lane >> 2, c = lane & 3;
    __syncwarp();  // every lane's ldmatrix reads of these rows are done (they are, behind the mma.sync chain; this says so)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        *reinterpret_cast<uint32_t*>(stage + off64(row0 + g, 8 * j + 2 * c)) = pack_bf16(o[j][0] * r0, o[j][1] * r0);
        *reinterpret_cast<uint32_t*>(stage + off64(row0 + g + 8, 8 * j + 2 * c)) = pack_bf16(o[j][2] * r1, o[j][3] * r1);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = row0 + 4 * i + (lane >> 3), ch = lane & 7;
        const uint4 val = *reinterpret_cast<const uint4*>(stage + off64(r, 8 * ch));
        *reinterpret_cast<uint4*>(gbase + (int64_t)r * g_st + 8 * ch) = val;
    }
}

// ------------------------------------------------------------------ forward
// Two (sequence, head) pairs are in flight per block: the q / k / v tiles of the next pair arrive (cp.async) while the
// current one is computed.  FULL: T == 128, no tile-range predicates.
template <bool FULL>
__global__ void __launch_bounds__(kThreads, 2) attention_fwd_kernel(const Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const uint32_t step = p.step + (p.step_ptr ? __ldg(p.step_ptr) : 0u);
    const int T = FULL ? TMAX : p.T;
    const int total = p.B * p.H;
    auto issue = [&](int pair, int buf) {
        const int b = pair / p.H, h = pair - b * p.H;
        uint8_t* const base = smem + buf * FWD_BUF;
        load_tile(base, p.q + b * p.q_sb + h * p.q_sh, p.q_st, T);
        load_tile(base + TILE_BYTES, p.k + b * p.k_sb + h * p.k_sh, p.k_st, T);
        load_tile(base + 2 * TILE_BYTES, p.v + b * p.v_sb + h * p.v_sh, p.v_st, T);
    };
    if ((int)blockIdx.x < total) issue(blockIdx.x, 0);
    cp_async_commit();
    int it = 0;
    for (int pair = blockIdx.x; pair < total; pair += gridDim.x, ++it) {
        const int b = pair / p.H, h = pair - b * p.H;
        if (pair + (int)gridDim.x < total) issue(pair + gridDim.x, (it + 1) & 1);
        cp_async_commit();
        cp_async_wait_group<1>();
        __syncthreads();
        uint8_t* const sQ = smem + (it & 1) * FWD_BUF;
        const uint8_t* const sK = sQ + TILE_BYTES;
        const uint8_t* const sV = sQ + 2 * TILE_BYTES;
        const int row0 = warp * 16;
        if (FULL || row0 < T) {
            float s[16][4];
            scores<8>(s, sQ, sK, row0, 0, lane);
            // ---- softmax of rows (row0 + g) and (row0 + g + 8); a row lives in the 4 lanes of a quad
            float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (FULL || 8 * j < T) {  // T % 16 == 0: a tile is either wholly inside or wholly outside
                    mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
                    mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
                mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            }
            float sum[2] = {0.0f, 0.0f};
            const float off0 = mx[0] * p.scale_log2e, off1 = mx[1] * p.scale_log2e;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const bool in = FULL || 8 * j < T;
                s[j][0] = in ? bf_ex2_approx(fmaf(s[j][0], p.scale_log2e, -off0)) : 0.0f;
                s[j][1] = in ? bf_ex2_approx(fmaf(s[j][1], p.scale_log2e, -off0)) : 0.0f;
                s[j][2] = in ? bf_ex2_approx(fmaf(s[j][2], p.scale_log2e, -off1)) : 0.0f;
                s[j][3] = in ? bf_ex2_approx(fmaf(s[j][3], p.scale_log2e, -off1)) : 0.0f;
                sum[0] += s[j][0] + s[j][1];
                sum[1] += s[j][2] + s[j][3];
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
                sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
            }
            const int64_t grow = ((int64_t)b * p.H + h) * T + row0 + g;  // global index of row (row0 + g)
            if (c == 0) {
                p.lse[grow] = off0 + bf_lg2_approx(sum[0]);
                p.lse[grow + 8] = off1 + bf_lg2_approx(sum[1]);
            }
            // ---- dropout on the (unnormalised) probabilities; 1 / sum is applied to the output rows
            if (p.thresh != 0u) {
                const uint32_t kb0 = keep_bits<4>(p, (uint32_t)grow, c, step, 0);
                const uint32_t kb1 = keep_bits<4>(p, (uint32_t)(grow + 8), c, step, 0);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    s[j][0] = ((kb0 >> (2 * j)) & 1u) ? s[j][0] * p.inv_keep : 0.0f;
                    s[j][1] = ((kb0 >> (2 * j + 1)) & 1u) ? s[j][1] * p.inv_keep : 0.0f;
                    s[j][2] = ((kb1 >> (2 * j)) & 1u) ? s[j][2] * p.inv_keep : 0.0f;
                    s[j][3] = ((kb1 >> (2 * j + 1)) & 1u) ? s[j][3] * p.inv_keep : 0.0f;
                }
            }
            // ---- O = P_d v: the score accumulators are already in A-fragment order
            float o[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.0f;
            const uint32_t v_base = smem_u32(sV);
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) {
                const uint32_t a[4] = {pack_bf16(s[2 * kc][0], s[2 * kc][1]), pack_bf16(s[2 * kc][2], s[2 * kc][3]),
                                       pack_bf16(s[2 * kc + 1][0], s[2 * kc + 1][1]),
                                       pack_bf16(s[2 * kc + 1][2], s[2 * kc + 1][3])};
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    uint32_t vb[4];
                    ldsm_x4_t(vb, v_base + off64(kc * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, np * 16 + (lane >> 4) * 8));
                    mma16816(o[2 * np], a, vb[0], vb[1]);
                    mma16816(o[2 * np + 1], a, vb[2], vb[3]);
                }
            }
            // out[b][t][h][:] through the warp's own (now dead) q rows: whole 128 B rows per store
            store_rows64(sQ, row0, o, bf_rcp_approx(sum[0]), bf_rcp_approx(sum[1]),
                         p.out + ((int64_t)b * T * p.H + h) * D, (int64_t)p.H * D, lane);
        }
        __syncthreads();  // the loads issued by the next iteration overwrite this buffer's partner... and then this one
    }
    cp_async_wait_group<0>();
}

// ------------------------------------------------------------------ backward
// 16 warps: warp w owns query rows 16 * (w % 8) and, in phase A, the key half w / 8 -- so a thread holds 32 scores,
// not 64, and four warps per scheduler hide each other's latencies.  The q / k / v / dO tiles of the next pair arrive
// while this one is computed.
//   phase A : S, P (from the stored log-sum-exp), keep mask, dP = dO v^T for (16 rows) x (64 keys);
//             partial D_i = sum_k P_d[i][k] dP_d[i][k] (== rowsum(dO o O), so O is never read) -> shared memory;
//             P_d -> sP.                       barrier
//             dS = P o (dP_d - D) * scale -> sS.   barrier
//   phase B : warps 0..7: dv (their 16 keys) = P_d^T dO;  warps 8..15: dk (their 16 keys) = dS^T q;
//             every warp: dq (its 16 rows, d half w / 8) = dS k.
template <bool FULL>
__global__ void __launch_bounds__(kBwdThreads, 1) attention_bwd_kernel(const Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* const sP = smem + 2 * BWD_BUF;  // P_d [q][key]
    uint8_t* const sS = sP + SQ_BYTES;       // dS  [q][key]
    float* const sD = reinterpret_cast<float*>(sS + SQ_BYTES);  // [2][128] partial D of the two key halves
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const int rw = warp & 7, kh = warp >> 3, row0 = rw * 16, key0 = 64 * kh;
    const uint32_t step = p.step + (p.step_ptr ? __ldg(p.step_ptr) : 0u);
    const int T = FULL ? TMAX : p.T;
    const int total = p.B * p.H;
    const int64_t o_st = (int64_t)p.H * D;
    auto issue = [&](int pair, int buf) {
        const int b = pair / p.H, h = pair - b * p.H;
        uint8_t* const base = smem + buf * BWD_BUF;
        load_tile(base, p.q + b * p.q_sb + h * p.q_sh, p.q_st, T);
        load_tile(base + TILE_BYTES, p.k + b * p.k_sb + h * p.k_sh, p.k_st, T);
        load_tile(base + 2 * TILE_BYTES, p.v + b * p.v_sb + h * p.v_sh, p.v_st, T);
        load_tile(base + 3 * TILE_BYTES, p.dout + ((int64_t)b * T * p.H + h) * D, o_st, T);
    };
    if ((int)blockIdx.x < total) issue(blockIdx.x, 0);
    cp_async_commit();
    const float ik = p.thresh != 0u ? p.inv_keep : 1.0f;
    int it = 0;
    for (int pair = blockIdx.x; pair < total; pair += gridDim.x, ++it) {
        const int b = pair / p.H, h = pair - b * p.H;
        const int64_t o_base = ((int64_t)b * T * p.H + h) * D;  // [b][0][h][0] of the [B, T, H, 64] tensors
        if (pair + (int)gridDim.x < total) issue(pair + gridDim.x, (it + 1) & 1);
        cp_async_commit();
        const bool rows_in = FULL || row0 < T;
        const int64_t grow = ((int64_t)b * p.H + h) * T + row0 + g;
        float lse0 = 0.0f, lse1 = 0.0f;
        if (rows_in) lse0 = __ldg(p.lse + grow), lse1 = __ldg(p.lse + grow + 8);
        cp_async_wait_group<1>();
        __syncthreads();
        const uint8_t* const sQ = smem + (it & 1) * BWD_BUF;
        const uint8_t* const sK = sQ + TILE_BYTES;
        const uint8_t* const sV = sQ + 2 * TILE_BYTES;
        const uint8_t* const sO = sQ + 3 * TILE_BYTES;  // dO

        // =============== phase A: 16 query rows x 64 keys per warp ===============
        float s[8][4], dp[8][4];
        uint32_t kb0 = 0xffffffffu, kb1 = 0xffffffffu;
        if (rows_in) {
            scores<4>(s, sQ, sK, row0, key0, lane);
            if (p.thresh != 0u) {
                kb0 = keep_bits<2>(p, (uint32_t)grow, c, step, 2 * kh);
                kb1 = keep_bits<2>(p, (uint32_t)(grow + 8), c, step, 2 * kh);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // true probabilities P
                const bool in = FULL || key0 + 8 * j < T;
                s[j][0] = in ? bf_ex2_approx(fmaf(s[j][0], p.scale_log2e, -lse0)) : 0.0f;
                s[j][1] = in ? bf_ex2_approx(fmaf(s[j][1], p.scale_log2e, -lse0)) : 0.0f;
                s[j][2] = in ? bf_ex2_approx(fmaf(s[j][2], p.scale_log2e, -lse1)) : 0.0f;
                s[j][3] = in ? bf_ex2_approx(fmaf(s[j][3], p.scale_log2e, -lse1)) : 0.0f;
            }
            // dP_d = dO_w v^T (same shape of product as the scores: A = dO rows, B = v as [key][d])
            scores<4>(dp, sO, sV, row0, key0, lane);
            float part0 = 0.0f, part1 = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                // dp <- keep ? dP_d / (1 - p) : 0 (the gradient w.r.t. the undropped probability)
                dp[j][0] = ((kb0 >> (2 * j)) & 1u) ? dp[j][0] * ik : 0.0f;
                dp[j][1] = ((kb0 >> (2 * j + 1)) & 1u) ? dp[j][1] * ik : 0.0f;
                dp[j][2] = ((kb1 >> (2 * j)) & 1u) ? dp[j][2] * ik : 0.0f;
                dp[j][3] = ((kb1 >> (2 * j + 1)) & 1u) ? dp[j][3] * ik : 0.0f;
                part0 = fmaf(s[j][0], dp[j][0], fmaf(s[j][1], dp[j][1], part0));
                part1 = fmaf(s[j][2], dp[j][2], fmaf(s[j][3], dp[j][3], part1));
            }
            part0 += __shfl_xor_sync(0xffffffffu, part0, 1);
            part0 += __shfl_xor_sync(0xffffffffu, part0, 2);
            part1 += __shfl_xor_sync(0xffffffffu, part1, 1);
            part1 += __shfl_xor_sync(0xffffffffu, part1, 2);
            if (c == 0) sD[kh * TMAX + row0 + g] = part0, sD[kh * TMAX + row0 + g + 8] = part1;
            // P_d as bf16 -> sP[q][key]
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
                const int j0 = 2 * kc, j1 = 2 * kc + 1, col = key0 + 16 * kc + 2 * c;
                *reinterpret_cast<uint32_t*>(sP + off128(row0 + g, col)) =
                    pack_bf16(((kb0 >> (2 * j0)) & 1u) ? s[j0][0] * ik : 0.0f, ((kb0 >> (2 * j0 + 1)) & 1u) ? s[j0][1] * ik : 0.0f);
                *reinterpret_cast<uint32_t*>(sP + off128(row0 + g + 8, col)) =
                    pack_bf16(((kb1 >> (2 * j0)) & 1u) ? s[j0][2] * ik : 0.0f, ((kb1 >> (2 * j0 + 1)) & 1u) ? s[j0][3] * ik : 0.0f);
                *reinterpret_cast<uint32_t*>(sP + off128(row0 + g, col + 8)) =
                    pack_bf16(((kb0 >> (2 * j1)) & 1u) ? s[j1][0] * ik : 0.0f, ((kb0 >> (2 * j1 + 1)) & 1u) ? s[j1][1] * ik : 0.0f);
                *reinterpret_cast<uint32_t*>(sP + off128(row0 + g + 8, col + 8)) =
                    pack_bf16(((kb1 >> (2 * j1)) & 1u) ? s[j1][2] * ik : 0.0f, ((kb1 >> (2 * j1 + 1)) & 1u) ? s[j1][3] * ik : 0.0f);
            }
        }
        __syncthreads();  // both halves of D are there
        if (rows_in) {
            // D_i = sum_k P[i][k] * (keep ? dP_d / (1-p) : 0) = sum_k P_d[i][k] dP_d[i][k]
            const float D0 = sD[row0 + g] + sD[TMAX + row0 + g], D1 = sD[row0 + g + 8] + sD[TMAX + row0 + g + 8];
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
                const int j0 = 2 * kc, j1 = 2 * kc + 1, col = key0 + 16 * kc + 2 * c;
                *reinterpret_cast<uint32_t*>(sS + off128(row0 + g, col)) =
                    pack_bf16(s[j0][0] * (dp[j0][0] - D0) * p.scale, s[j0][1] * (dp[j0][1] - D0) * p.scale);
                *reinterpret_cast<uint32_t*>(sS + off128(row0 + g + 8, col)) =
                    pack_bf16(s[j0][2] * (dp[j0][2] - D1) * p.scale, s[j0][3] * (dp[j0][3] - D1) * p.scale);
                *reinterpret_cast<uint32_t*>(sS + off128(row0 + g, col + 8)) =
                    pack_bf16(s[j1][0] * (dp[j1][0] - D0) * p.scale, s[j1][1] * (dp[j1][1] - D0) * p.scale);
                *reinterpret_cast<uint32_t*>(sS + off128(row0 + g + 8, col + 8)) =
                    pack_bf16(s[j1][2] * (dp[j1][2] - D1) * p.scale, s[j1][3] * (dp[j1][3] - D1) * p.scale);
            }
        }
        __syncthreads();  // P_d and dS complete

        // =============== phase B ===============
        if (rows_in) {
            // ---- dv (warps 0..7) or dk (warps 8..15) of the 16 keys row0 .. row0 + 15: A = transposed 16 x 16 blocks of
            //      sP / sS, B = dO / q as [query][d]
            float acc[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;
            const uint32_t a_base = smem_u32(kh == 0 ? sP : sS), b_base = smem_u32(kh == 0 ? sO : sQ);
            const int mi = lane >> 3;
#pragma unroll
            for (int qc = 0; qc < 8; ++qc) {
                if (!FULL && 16 * qc >= T) break;
                uint32_t a[4];
                ldsm_x4_t(a, a_base + off128(qc * 16 + (lane & 7) + (mi >> 1) * 8, row0 + (mi & 1) * 8));
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    uint32_t bb[4];
                    ldsm_x4_t(bb, b_base + off64(qc * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, np * 16 + (lane >> 4) * 8));
                    mma16816(acc[2 * np], a, bb[0], bb[1]);
                    mma16816(acc[2 * np + 1], a, bb[2], bb[3]);
                }
            }
            __nv_bfloat16* const orow = (kh == 0 ? p.dv : p.dk) + o_base + (int64_t)(row0 + g) * o_st + 2 * c;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                *reinterpret_cast<uint32_t*>(orow + 8 * j) = pack_bf16(acc[j][0], acc[j][1]);
                *reinterpret_cast<uint32_t*>(orow + 8 * o_st + 8 * j) = pack_bf16(acc[j][2], acc[j][3]);
            }
            // ---- dq of the 16 rows, columns 32 * kh .. +31: A = sS rows, B = k as [key][d] read transposed
            float dq[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.0f;
            const uint32_t s_base = smem_u32(sS), k_base = smem_u32(sK);
#pragma unroll
            for (int kc = 0; kc < 8; ++kc) {
                if (!FULL && 16 * kc >= T) break;
                uint32_t a[4];
                ldsm_x4(a, s_base + off128(row0 + (lane & 7) + ((lane >> 3) & 1) * 8, kc * 16 + (lane >> 4) * 8));
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    uint32_t kb[4];
                    ldsm_x4_t(kb, k_base + off64(kc * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, (2 * kh + np) * 16 + (lane >> 4) * 8));
                    mma16816(dq[2 * np], a, kb[0], kb[1]);
                    mma16816(dq[2 * np + 1], a, kb[2], kb[3]);
                }
            }
            __nv_bfloat16* const qrow = p.dq + o_base + (int64_t)(row0 + g) * o_st + 32 * kh + 2 * c;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                *reinterpret_cast<uint32_t*>(qrow + 8 * j) = pack_bf16(dq[j][0], dq[j][1]);
                *reinterpret_cast<uint32_t*>(qrow + 8 * o_st + 8 * j) = pack_bf16(dq[j][2], dq[j][3]);
            }
        }
        __syncthreads();  // sP / sS / sD and this tile buffer are free again
    }
    cp_async_wait_group<0>();
}

__global__ void __launch_bounds__(256) attention_mask_kernel(uint8_t* out, const Params p) {
    const uint32_t step = p.step + (p.step_ptr ? __ldg(p.step_ptr) : 0u);
    const int64_t n = (int64_t)p.B * p.H * p.T * p.T;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int k = (int)(i % p.T);
        const uint32_t row = (uint32_t)(i / p.T);  // (b * H + h) * T + q
        uint8_t keep = 1;
        if (p.thresh != 0u) {
            const uint4 r = bf_philox4x32_10(row, (uint32_t)((k % 8) / 2 + 4 * (k / 32)), 0x40000000u | p.site, step, p.k0, p.k1);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
            const uint32_t word = w[(k / 8) % 4];
            const uint32_t u = (k & 1) ? (word >> 16) : (word & 0xffffu);
            keep = u >= p.thresh;
        }
        out[i] = keep;
    }
}

static int fill_common(Params& p, int64_t B, int64_t H, int64_t T, float scale, float p_drop, uint64_t seed, uint32_t step,
                       uint32_t site) {
    p.B = (int)B, p.H = (int)H, p.T = (int)T;
    p.scale = scale;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.thresh = 0u;
    p.inv_keep = 1.0f;
    if (p_drop > 0.0f) {
        uint32_t t = (uint32_t)lrintf(p_drop * 65536.0f);
        if (t > 65535u) t = 65535u;
        p.thresh = t;
        p.inv_keep = 65536.0f / (65536.0f - (float)t);  // 1 / P(keep) of the quantised mask
    }
    p.k0 = (uint32_t)(seed & 0xffffffffu), p.k1 = (uint32_t)(seed >> 32), p.step = step, p.site = site;
    p.step_ptr = bf_step_counter();
    return 0;
}

}  // namespace attn

#define BF_ATTN_CHECK()                                                                                         \
    BF_CHECK_ARG(B >= 1 && H >= 1 && T >= 16 && T <= attn::TMAX && T % 16 == 0, "needs 16 <= T <= 128, T % 16 == 0"); \
    BF_CHECK_ARG(p_drop >= 0.0f && p_drop < 1.0f, "bad dropout probability");                                      \
    BF_CHECK_ARG(B * H * T < (int64_t)1 << 32, "too many rows for the mask counter")

// bf_attention_tc.cu
int bf_attention_tc_fwd(const void* q, const void* k, const void* v, const int64_t* strides, int64_t B, int64_t H, float scale,
                        float p_drop, uint64_t seed, uint32_t step, uint32_t site, void* out, float* lse, uint32_t* keep,
                        cudaStream_t stream);
int bf_attention_tc_bwd(const void* dout, const void* q, const void* k, const void* v, const int64_t* strides, const float* lse,
                        const uint32_t* keep, int64_t B, int64_t H, float scale, float p_drop, uint64_t seed, uint32_t step,
                        uint32_t site, void* dq, void* dk, void* dv, float* dbias, void* workspace, int64_t S,
                        cudaStream_t stream);
int64_t bf_attention_tc_bias_workspace_bytes(int64_t B, int64_t H, int64_t S);

extern "C" int bf_attention_supported(int64_t T, int64_t head_dim) {
    return (head_dim == attn::D && T >= 16 && T <= attn::TMAX && T % 16 == 0) ? 1 : 0;
}

extern "C" int bf_attention_fwd(const void* q, const void* k, const void* v, const int64_t* strides, int64_t B, int64_t H,
                                int64_t T, float scale, float p_drop, uint64_t seed, uint32_t step, uint32_t site,
                                void* out, float* lse, uint32_t* keep, void* stream) {
    BF_CHECK_ARG(q && k && v && strides && out && lse, "null pointer");
    BF_ATTN_CHECK();
    for (int i = 0; i < 9; ++i) BF_CHECK_ARG(strides[i] % 8 == 0, "strides must be multiples of 8 elements (16 B)");
    BF_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                   reinterpret_cast<uintptr_t>(out)) & 15u) == 0, "q, k, v, out must be 16 B aligned");
    if (T == attn::TMAX && bf_option(BF_OPT_ATTN_TC) && (keep || p_drop == 0.0f))
        return bf_attention_tc_fwd(q, k, v, strides, B, H, scale, p_drop, seed, step, site, out, lse, keep,
                                   reinterpret_cast<cudaStream_t>(stream));
    attn::Params p{};
    p.q = (const __nv_bfloat16*)q, p.k = (const __nv_bfloat16*)k, p.v = (const __nv_bfloat16*)v;
    p.out = (__nv_bfloat16*)out, p.lse = lse;
    p.q_sb = strides[0], p.q_sh = strides[1], p.q_st = strides[2];
    p.k_sb = strides[3], p.k_sh = strides[4], p.k_st = strides[5];
    p.v_sb = strides[6], p.v_sh = strides[7], p.v_st = strides[8];
    attn::fill_common(p, B, H, T, scale, p_drop, seed, step, site);
    auto* const kernel = T == attn::TMAX ? attn::attention_fwd_kernel<true> : attn::attention_fwd_kernel<false>;
    BF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::FWD_SMEM));
    const int64_t pairs = B * H, cap = (int64_t)bf_num_sms() * 2;
    kernel<<<(int)(pairs < cap ? pairs : cap), attn::kThreads, attn::FWD_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    BF_LAUNCH_OK();
    return 0;
}

extern "C" int bf_attention_bwd_bias(const void* dout, const void* q, const void* k, const void* v, const int64_t* strides,
                                const void* out, const float* lse, const uint32_t* keep, int64_t B, int64_t H, int64_t T,
                                float scale, float p_drop, uint64_t seed, uint32_t step, uint32_t site, void* dq, void* dk,
                                void* dv, float* dbias, void* workspace, int64_t S, void* stream) {
    BF_CHECK_ARG(dout && q && k && v && strides && lse && dq && dk && dv, "null pointer");
    BF_ATTN_CHECK();
    for (int i = 0; i < 9; ++i) BF_CHECK_ARG(strides[i] % 8 == 0, "strides must be multiples of 8 elements (16 B)");
    BF_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                   reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) |
                   reinterpret_cast<uintptr_t>(dv)) & 15u) == 0, "buffers must be 16 B aligned");
    const bool tc_path = T == attn::TMAX && bf_option(BF_OPT_ATTN_TC) && (keep || p_drop == 0.0f);
    BF_CHECK_ARG(!dbias || (tc_path && workspace && S >= 1 && B % S == 0),
                 "the fused q / k / v bias gradients need the tcgen05 path (T == 128, keep bits), a workspace and B % S == 0");
    if (tc_path)
        return bf_attention_tc_bwd(dout, q, k, v, strides, lse, keep, B, H, scale, p_drop, seed, step, site, dq, dk, dv, dbias,
                                   workspace, S, reinterpret_cast<cudaStream_t>(stream));
    attn::Params p{};
    p.q = (const __nv_bfloat16*)q, p.k = (const __nv_bfloat16*)k, p.v = (const __nv_bfloat16*)v;
    p.o = (const __nv_bfloat16*)out, p.dout = (const __nv_bfloat16*)dout, p.lse = const_cast<float*>(lse);
    p.dq = (__nv_bfloat16*)dq, p.dk = (__nv_bfloat16*)dk, p.dv = (__nv_bfloat16*)dv;
    p.q_sb = strides[0], p.q_sh = strides[1], p.q_st = strides[2];
    p.k_sb = strides[3], p.k_sh = strides[4], p.k_st = strides[5];
    p.v_sb = strides[6], p.v_sh = strides[7], p.v_st = strides[8];
    attn::fill_common(p, B, H, T, scale, p_drop, seed, step, site);
    auto* const kernel = T == attn::TMAX ? attn::attention_bwd_kernel<true> : attn::attention_bwd_kernel<false>;
    BF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::BWD_SMEM));
    const int64_t pairs = B * H, cap = (int64_t)bf_num_sms();
    kernel<<<(int)(pairs < cap ? pairs : cap), attn::kBwdThreads, attn::BWD_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    BF_LAUNCH_OK();
    return 0;
}

extern "C" int bf_attention_bwd(const void* dout, const void* q, const void* k, const void* v, const int64_t* strides,
                                const void* out, const float* lse, const uint32_t* keep, int64_t B, int64_t H, int64_t T,
                                float scale, float p_drop, uint64_t seed, uint32_t step, uint32_t site, void* dq, void* dk,
                                void* dv, void* stream) {
    return bf_attention_bwd_bias(dout, q, k, v, strides, out, lse, keep, B, H, T, scale, p_drop, seed, step, site, dq, dk, dv,
                                 nullptr, nullptr, 1, stream);
}

extern "C" int64_t bf_attention_bias_workspace_bytes(int64_t B, int64_t H, int64_t S) {
    return bf_attention_tc_bias_workspace_bytes(B < 1 ? 1 : B, H < 1 ? 1 : H, S < 1 ? 1 : S);
}

extern "C" int bf_attention_dropout_mask(uint8_t* out, int64_t B, int64_t H, int64_t T, float p_drop, uint64_t seed,
                                         uint32_t step, uint32_t site, void* stream) {
    BF_CHECK_ARG(out, "null pointer");
    BF_ATTN_CHECK();
    attn::Params p{};
    attn::fill_common(p, B, H, T, 1.0f, p_drop, seed, step, site);
    attn::attention_mask_kernel<<<bf_num_sms() * 4, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, p);
    BF_LAUNCH_OK();
    return 0;
}
