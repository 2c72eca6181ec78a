// tcgen05 version of the short-sequence attention (T == 128, head width 64): the five products of a (sequence, head)
// pair are whole-tile tensor-core MMAs on TMA-loaded shared-memory operands with TMEM accumulators, so the CUDA cores
// only do what is left -- the softmax, the dropout mask and the packing -- on ONE row per thread.  bf_attention.cu
// (mma.sync on ldmatrix fragments) stays for the other lengths; both produce the same keep mask.
//
// One block per SM, 18 warps:
//   warps 0..15 row workers: thread (w, lane) owns row 32 * (w % 4) + lane of every accumulator (its TMEM lane) and the
//               column quarter w / 4 -- 32 scores, or 16 output columns
//   warp 16     TMA producer (one thread): q, k, v (and dO) tiles of the NEXT pair while this one is computed
//   warp 17     MMA issuer (one thread)
//   forward: warps 18..25 generate the Philox keep bits of the pairs ahead (they fill the issue slots the row workers
//   leave idle; in the row workers the mask was 2/3 of all instructions); backward: warp 18 stores the gradient tiles
// Forward outputs leave straight from registers: a thread holds 16 consecutive bf16 of an output row = one 32 B sector
// (st.global.v8).  The backward stages its three gradient tiles where P_d / dS were and a store warp sends them off by
// TMA while the next pair's second pass computes (direct stores of 48 KB per pair were LSU-bound).
//
// forward, per pair:   S = q k^T  (128x128x64, TMEM)        -> rows: max / exp2 / sum (four quarters exchanged through
//                      shared memory), Philox keep bits (stored: 16 B per row, the backward does not regenerate them),
//                      P_d as bf16 -> shared memory          -> O = P_d v (128x64x128)   -> rows: * 1/((1-p) sum) ->
//                      global memory.  S, P_d and O are double buffered: S(i+1) and O(i-1) run under the softmax of i.
// backward, per pair:  S = q k^T, dP = dO v^T                -> rows: P = exp2(S c - lse), D = sum_k P_d dP_d (== dO.O:
//                      O is never read), P_d and dS as bf16 -> shared memory [q][key], which the tensor core reads
//                      K-major for dq = dS k and MN-major (transposed) for dv = P_d^T dO and dk = dS^T q
//                      -> rows -> global memory.  The first pass of pair i+1 runs under the dv / dk / dq MMAs of pair i.
// Every output element is written once; nothing is accumulated across blocks (deterministic).
#include "bf_tc.cuh"

namespace attn_tc {
using namespace tc;

// phase timeline of block 0 (scripts/attn_trace.cu builds this file with -DBF_ATTN_TRACE); compiled out otherwise
#ifdef BF_ATTN_TRACE
__device__ unsigned long long g_trace[2 * 16 * 16];
#define BF_STAMP(who, slot) \
    do { if (blockIdx.x == 0 && i < 16) g_trace[((who) * 16 + i) * 16 + (slot)] = clock64(); } while (0)
#else
#define BF_STAMP(who, slot) do { } while (0)
#endif

constexpr int TT = 128, DD = 64;
constexpr int TILE = TT * DD * 2;      // one [128][64] bf16 tile, 128 B rows, 128B-swizzled: 16 KiB
constexpr int kRowWarps = 16;
constexpr int kRowThreads = kRowWarps * 32;   // 512 row workers
constexpr int kThreads = kRowThreads + 64;    // + producer warp (16) + MMA warp (17)
constexpr int kBwdThreads = kThreads + 64;    // backward: + store / column-sum warps (18, 19)
constexpr int kMaskThreads = 256;             // forward: + 8 mask warps (18..25), two threads per query row
constexpr int kFwdThreads = kThreads + kMaskThreads;
constexpr int KEEP_BYTES = 2 * TT * 16;       // forward: keep words of two pairs
constexpr int FWD_BUF = 3 * TILE;      // q, k, v
constexpr int BWD_BUF = 4 * TILE;      // q, k, v, dO
constexpr int XCH_BYTES = 2 * 4 * TT * 4;  // two exchanged row statistics x four column quarters
constexpr int BAR_BYTES = 256;
constexpr int kFwdStages = 3;          // q, k, v tile buffers in flight (the kernel is bound by bytes in flight per SM)
constexpr int FWD_SMEM = 1024 + kFwdStages * FWD_BUF + 4 * TILE + XCH_BYTES + KEEP_BYTES + BAR_BYTES;  // + 2 x P_d (2 sub-tiles)
constexpr int BWD_SMEM = 1024 + 2 * BWD_BUF + 4 * TILE + XCH_BYTES + BAR_BYTES;         // + P_d, dS (2 sub-tiles each)
constexpr uint32_t TMEM_COLS = 512;

struct Params {
    __nv_bfloat16* out;  // forward: [B, 128, heads, 64], written row quarter by row quarter (32 B per thread); the
                         // backward's dq / dk / dv leave through TMA stores (tensor maps)
    float* lse;       // [B, heads, 128] base-2 log-sum-exp of the scaled scores
    uint32_t* keep;   // [B, heads, 128, 4] keep bits of the 128 keys of a row (bit k % 32 of word k / 32), or null
    int H, total;     // heads, B * heads
    // backward only: per-block partial column sums of dq, dk, dv -- the bias gradients of the q / k / v projections --
    // [gridDim.x][3][S][heads * 64] fp32 (null: not wanted); seqs_per_sample = B / S (the batch is S samples folded)
    float* col_partial;
    int S, seqs_per_sample;
    float scale, scale_log2e, inv_keep;
    uint32_t thresh, k0, k1, step, site;
    const uint32_t* step_ptr;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns (see tmem_ld_32x32)
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// 32 B (16 bf16 = one output row quarter, one full sector) straight from registers
__device__ __forceinline__ void stg256(void* ptr, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}
// 16 B chunk `chunk` of row `row` in a swizzled [128][64] bf16 tile
__device__ __forceinline__ uint32_t chunk_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }
// MN-major operand: rows = reduction index (128 B each, 8-row groups 1024 B apart), 64-wide atoms one tile apart
__device__ __forceinline__ uint64_t mn_desc(uint32_t tile, int k) { return make_smem_desc(tile + k * 2048, TILE, 1024); }
// K-major operand with a 128-deep reduction: reduction atoms (64 elements) one tile apart
__device__ __forceinline__ uint64_t k_desc128(uint32_t tile, int k) { return operand_desc<false>(tile + (k >> 2) * TILE, k & 3); }

// keep bits of the 32 keys [32 cq, 32 cq + 32) of query row `row` (bit k % 32): the mask bf_attention.cu defines --
// Philox call (row, (k % 8) / 2 + 4 * (k / 32)), word (k / 8) % 4, half k % 2
__device__ __forceinline__ uint32_t keep_word(const Params& p, uint32_t row, int cq, uint32_t step) {
    uint32_t w = 0u;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint4 r = bf_philox4x32_10(row, (uint32_t)(c + 4 * cq), 0x40000000u | p.site, step, p.k0, p.k1);
        const uint32_t v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            w |= (uint32_t)((v[t] & 0xffffu) >= p.thresh) << (8 * t + 2 * c);
            w |= (uint32_t)((v[t] >> 16) >= p.thresh) << (8 * t + 2 * c + 1);
        }
    }
    return w;
}

struct Smem {
    uint32_t base;   // 1024 B aligned shared-window address
    uint8_t* gen;    // the same place as a generic pointer
};
__device__ __forceinline__ Smem aligned_smem(uint8_t* raw) {
    const uint32_t a = smem_u32(raw), b = (a + 1023u) & ~1023u;
    return {b, raw + (b - a)};
}

// Row workers: thread (w, lane) owns row 32 * (w % 4) + lane (its TMEM lane) and the column quarter w / 4:
// 32 scores, or 16 output columns.

// ------------------------------------------------------------------ forward
// S, P_d and O are double buffered, so per pair the row workers only ever wait for S: while they do the softmax of
// pair i the tensor core runs O(i-1) = P_d v and S(i+1) = q k^T; the output rows of pair i-1 are drained after the
// softmax of pair i.
__global__ void __launch_bounds__(kFwdThreads, 1)
    fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_v, const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    const Smem sm = aligned_smem(smem_raw);
    const uint32_t s_tiles = sm.base, s_p = sm.base + kFwdStages * FWD_BUF;
    float* const xch = reinterpret_cast<float*>(sm.gen + kFwdStages * FWD_BUF + 4 * TILE);  // [kind][quarter][row]
    uint4* const keep_s = reinterpret_cast<uint4*>(sm.gen + kFwdStages * FWD_BUF + 4 * TILE + XCH_BYTES);  // [2][row]
    const uint32_t bars = s_p + 4 * TILE + XCH_BYTES + KEEP_BYTES;
    auto full = [&](int st) { return bars + 8u * st; };                 // q, k, v tiles: kFwdStages deep
    auto empty = [&](int st) { return bars + 8u * (kFwdStages + st); };
    auto s_ready = [&](int b) { return bars + 8u * (2 * kFwdStages + b); };  // S, P_d, O: double buffered (b = i % 2)
    auto s_free = [&](int b) { return bars + 8u * (2 * kFwdStages + 2 + b); };
    auto p_ready = [&](int b) { return bars + 8u * (2 * kFwdStages + 4 + b); };
    auto o_ready = [&](int b) { return bars + 8u * (2 * kFwdStages + 6 + b); };
    auto o_free = [&](int b) { return bars + 8u * (2 * kFwdStages + 8 + b); };
    auto mask_ready = [&](int b) { return bars + 8u * (2 * kFwdStages + 10 + b); };
    auto mask_free = [&](int b) { return bars + 8u * (2 * kFwdStages + 12 + b); };
    const uint32_t tmem_slot = bars + 160;
    volatile uint32_t* const tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(sm.gen + kFwdStages * FWD_BUF + 4 * TILE + XCH_BYTES + KEEP_BYTES + 160);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == kRowThreads) {
        tma_prefetch_desc(&map_q), tma_prefetch_desc(&map_k), tma_prefetch_desc(&map_v);
        for (int st = 0; st < kFwdStages; ++st) mbar_init(full(st), 1), mbar_init(empty(st), 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(s_ready(b), 1), mbar_init(s_free(b), kRowThreads);
            mbar_init(p_ready(b), kRowThreads), mbar_init(o_ready(b), 1), mbar_init(o_free(b), kRowThreads);
            mbar_init(mask_ready(b), kMaskThreads), mbar_init(mask_free(b), kRowThreads);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == kRowWarps + 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_gen;
    const int n_mine = p.total > (int)blockIdx.x ? (p.total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    constexpr uint32_t C_O = 2 * TT;  // TMEM columns: S buffers at 0 and 128, O buffers at 256 and 320

    if (warp == kRowWarps) {
        if (lane == 0) {
            for (int i = 0; i < n_mine; ++i) {
                const int pair = blockIdx.x + i * gridDim.x, st = i % kFwdStages, bb = pair / p.H, h = pair - bb * p.H;
                mbar_wait(empty(st), ((i / kFwdStages) & 1) ^ 1u);
                mbar_expect_tx(full(st), 3 * TILE);
                const uint32_t dst = s_tiles + st * FWD_BUF;
                tma_load_4d(dst, &map_q, full(st), 0, 0, h, bb);
                tma_load_4d(dst + TILE, &map_k, full(st), 0, 0, h, bb);
                tma_load_4d(dst + 2 * TILE, &map_v, full(st), 0, 0, h, bb);
            }
        }
    } else if (warp == kRowWarps + 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc(false, false, TT, TT), idesc_o = make_idesc(false, true, TT, DD);
            auto issue_s = [&](int i) {  // S(i) = q k^T into TMEM buffer i % 2
                const int b = i & 1, st = i % kFwdStages;
                mbar_wait(full(st), (i / kFwdStages) & 1);
                mbar_wait(s_free(b), ((i >> 1) & 1) ^ 1u);
                tc_fence_after();
                const uint32_t q = s_tiles + st * FWD_BUF, k = q + TILE;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tmem + b * TT, operand_desc<false>(q, kk), operand_desc<false>(k, kk), idesc_s, kk > 0);
                umma_commit(s_ready(b));
            };
            if (n_mine > 0) issue_s(0);
            for (int i = 0; i < n_mine; ++i) {
                const int b = i & 1;
                if (i + 1 < n_mine) issue_s(i + 1);  // runs under the softmax of pair i
                mbar_wait(p_ready(b), (i >> 1) & 1);
                mbar_wait(o_free(b), ((i >> 1) & 1) ^ 1u);
                tc_fence_after();
                const uint32_t v = s_tiles + (i % kFwdStages) * FWD_BUF + 2 * TILE;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)  // O = P_d v: reduction over the 128 keys
                    umma_bf16(tmem + C_O + b * DD, k_desc128(s_p + b * 2 * TILE, kk), mn_desc(v, kk), idesc_o, kk > 0);
                umma_commit(o_ready(b));
                umma_commit(empty(i % kFwdStages));  // q, k, v of this pair are no longer read
            }
        }
    } else if (warp >= kRowWarps + 2) {
        // ===================== mask warps: the Philox keep bits of the pairs ahead, off the row workers' critical path
        // (thread t: query row t / 2, key half t % 2: 8 calls -> two keep words (64 keys) -> shared + global memory)
        if (p.thresh != 0u) {
            const int t = threadIdx.x - (kRowThreads + 64), row = t >> 1, half = t & 1;
            const uint32_t step = p.step + (p.step_ptr ? __ldg(p.step_ptr) : 0u);
            for (int i = 0; i < n_mine; ++i) {
                const int pair = blockIdx.x + i * gridDim.x, b = i & 1;
                const uint32_t grow = (uint32_t)pair * TT + row;
                const uint2 w = make_uint2(keep_word(p, grow, 2 * half, step), keep_word(p, grow, 2 * half + 1, step));
                if (p.keep) *reinterpret_cast<uint2*>(p.keep + (size_t)grow * 4 + 2 * half) = w;
                mbar_wait(mask_free(b), ((i >> 1) & 1) ^ 1u);
                reinterpret_cast<uint2*>(keep_s + b * TT + row)[half] = w;
                mbar_arrive(mask_ready(b));
            }
        }
    } else {
        const int q = 32 * (warp & 3) + lane, cq = warp >> 2;
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        float r_prev = 0.0f;
        // output row q of pair j, columns 16 cq .. +15: straight from TMEM to global memory
        auto drain = [&](int j, float r) {
            const int pb = j & 1, pair = blockIdx.x + j * gridDim.x, bb = pair / p.H, h = pair - bb * p.H;
            uint32_t ro[16], o[8];
            mbar_wait(o_ready(pb), (j >> 1) & 1);
            tc_fence_after();
            tmem_ld_32x16(lane_addr + C_O + pb * DD + 16 * cq, ro);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(o_free(pb));
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = pack_bf16(__uint_as_float(ro[2 * e]) * r, __uint_as_float(ro[2 * e + 1]) * r);
            stg256(p.out + (((size_t)bb * TT + q) * p.H + h) * DD + 16 * cq, o);
        };
        for (int i = 0; i < n_mine; ++i) {
            const int pair = blockIdx.x + i * gridDim.x, b = i & 1;
            const uint32_t grow = (uint32_t)pair * TT + q;
            uint32_t ra[32];
            if (threadIdx.x == 0) BF_STAMP(0, 0);
            mbar_wait(s_ready(b), (i >> 1) & 1);
            if (threadIdx.x == 0) BF_STAMP(0, 1);
            tc_fence_after();
            tmem_ld_32x32(lane_addr + b * TT + 32 * cq, ra);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_free(b));
            float s[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) s[j] = __uint_as_float(ra[j]);
            float mx = s[0];
#pragma unroll
            for (int j = 1; j < 32; ++j) mx = fmaxf(mx, s[j]);
            xch[cq * TT + q] = mx;
            if (threadIdx.x == 0) BF_STAMP(0, 2);
            named_bar_sync<1, kRowThreads>();
            if (threadIdx.x == 0) BF_STAMP(0, 3);
            if (threadIdx.x == 0) BF_STAMP(0, 4);
            const float off = fmaxf(fmaxf(xch[q], xch[TT + q]), fmaxf(xch[2 * TT + q], xch[3 * TT + q])) * p.scale_log2e;
            float sum = 0.0f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                s[j] = bf_ex2_approx(fmaf(s[j], p.scale_log2e, -off));
                sum += s[j];
            }
            xch[4 * TT + cq * TT + q] = sum;
            if (threadIdx.x == 0) BF_STAMP(0, 5);
            named_bar_sync<2, kRowThreads>();
            if (threadIdx.x == 0) BF_STAMP(0, 6);
            const float total = (xch[4 * TT + q] + xch[5 * TT + q]) + (xch[6 * TT + q] + xch[7 * TT + q]);
            if (cq == 0) p.lse[grow] = off + bf_lg2_approx(total);
            const float r_cur = p.inv_keep * bf_rcp_approx(total);
            if (p.thresh != 0u) {
                mbar_wait(mask_ready(b), (i >> 1) & 1);
                const uint32_t w = reinterpret_cast<const uint32_t*>(keep_s + b * TT + q)[cq];
                mbar_arrive(mask_free(b));
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (!(w & (1u << j))) s[j] = 0.0f;
            }
            // P_d (still without 1 / ((1-p) sum): applied to the output row) -> row q of key sub-tile cq / 2
            const uint32_t dst = s_p + (b * 2 + (cq >> 1)) * TILE;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                sts128(dst + chunk_off(q, 4 * (cq & 1) + c), pack_bf16(s[8 * c], s[8 * c + 1]), pack_bf16(s[8 * c + 2], s[8 * c + 3]),
                       pack_bf16(s[8 * c + 4], s[8 * c + 5]), pack_bf16(s[8 * c + 6], s[8 * c + 7]));
            fence_proxy_async();
            mbar_arrive(p_ready(b));
            if (threadIdx.x == 0) BF_STAMP(0, 7);
            if (i > 0) drain(i - 1, r_prev);  // O(i-1) finished under this pair's softmax
            r_prev = r_cur;
        }
        if (n_mine > 0) drain(n_mine - 1, r_prev);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kRowWarps + 1) {
        tc_fence_after();
        tmem_dealloc(tmem, TMEM_COLS);
    }
}

// ------------------------------------------------------------------ backward
// Software pipelined over pairs: the first pass of pair i+1 (P, t = keep ? dP_d : 0 and the partial D) runs while the
// tensor core computes dv, dk, dq of pair i, whose rows are drained afterwards; the second pass (P_d, dS -> shared
// memory) runs while S and dP of the next pair are formed.  lse and the keep words are fetched one pair ahead.
__global__ void __launch_bounds__(kBwdThreads, 1)
    bwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
               const __grid_constant__ CUtensorMap map_dq, const __grid_constant__ CUtensorMap map_dk,
               const __grid_constant__ CUtensorMap map_dv, const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    const Smem sm = aligned_smem(smem_raw);
    const uint32_t s_tiles = sm.base, s_p = sm.base + 2 * BWD_BUF, s_s = s_p + 2 * TILE;
    float* const xch = reinterpret_cast<float*>(sm.gen + 2 * BWD_BUF + 4 * TILE);  // [pair parity][quarter][row]
    const uint32_t bars = s_s + 2 * TILE + XCH_BYTES;
    auto full = [&](int b) { return bars + 8u * b; };
    auto empty = [&](int b) { return bars + 8u * (2 + b); };
    const uint32_t s_ready = bars + 32, s_free = bars + 40, p_ready = bars + 48, o_ready = bars + 56, o_free = bars + 64,
                   staged = bars + 72, stage_free = bars + 80, tmem_slot = bars + 128;
    volatile uint32_t* const tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sm.gen + 2 * BWD_BUF + 4 * TILE + XCH_BYTES + 128);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == kRowThreads) {
        tma_prefetch_desc(&map_q), tma_prefetch_desc(&map_k), tma_prefetch_desc(&map_v), tma_prefetch_desc(&map_do);
        tma_prefetch_desc(&map_dq), tma_prefetch_desc(&map_dk), tma_prefetch_desc(&map_dv);
        for (int b = 0; b < 2; ++b) mbar_init(full(b), 1), mbar_init(empty(b), 1);
        mbar_init(s_ready, 1), mbar_init(s_free, kRowThreads), mbar_init(p_ready, kRowThreads), mbar_init(o_ready, 1);
        mbar_init(o_free, kRowThreads), mbar_init(staged, kRowThreads), mbar_init(stage_free, 2);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == kRowWarps + 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_gen;
    const int n_mine = p.total > (int)blockIdx.x ? (p.total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    // TMEM columns: S 0..127, dP 128..255, dv 256..319, dk 320..383, dq 384..447
    constexpr uint32_t C_S = 0, C_DP = 128, C_DV = 256, C_DK = 320, C_DQ = 384;

    if (warp == kRowWarps) {
        if (lane == 0) {
            for (int i = 0; i < n_mine; ++i) {
                const int pair = blockIdx.x + i * gridDim.x, b = i & 1, bb = pair / p.H, h = pair - bb * p.H;
                mbar_wait(empty(b), ((i >> 1) & 1) ^ 1u);
                mbar_expect_tx(full(b), 4 * TILE);
                const uint32_t dst = s_tiles + b * BWD_BUF;
                tma_load_4d(dst, &map_q, full(b), 0, 0, h, bb);
                tma_load_4d(dst + TILE, &map_k, full(b), 0, 0, h, bb);
                tma_load_4d(dst + 2 * TILE, &map_v, full(b), 0, 0, h, bb);
                tma_load_4d(dst + 3 * TILE, &map_do, full(b), 0, 0, h, bb);
            }
        }
    } else if (warp == kRowWarps + 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc(false, false, TT, TT);   // S, dP: both operands K-major
            constexpr uint32_t idesc_t = make_idesc(true, true, TT, DD);     // dv, dk: transposed A, B = [q][d]
            constexpr uint32_t idesc_q = make_idesc(false, true, TT, DD);    // dq: A = dS [q][key], B = k [key][d]
            auto issue_sdp = [&](int i) {
                const int b = i & 1;
                const uint32_t q = s_tiles + b * BWD_BUF, k = q + TILE, v = q + 2 * TILE, dO = q + 3 * TILE;
                mbar_wait(full(b), (i >> 1) & 1);
                mbar_wait(s_free, (i & 1) ^ 1u);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tmem + C_S, operand_desc<false>(q, kk), operand_desc<false>(k, kk), idesc_s, kk > 0);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tmem + C_DP, operand_desc<false>(dO, kk), operand_desc<false>(v, kk), idesc_s, kk > 0);
                umma_commit(s_ready);
            };
            if (n_mine > 0) issue_sdp(0);
            for (int i = 0; i < n_mine; ++i) {
                const int b = i & 1;
                const uint32_t q = s_tiles + b * BWD_BUF, k = q + TILE, dO = q + 3 * TILE;
                if (i + 1 < n_mine) issue_sdp(i + 1);  // as soon as the row workers have taken S, dP of pair i
                BF_STAMP(1, 3);
                mbar_wait(p_ready, i & 1);
                BF_STAMP(1, 4);
                mbar_wait(o_free, (i & 1) ^ 1u);
                BF_STAMP(1, 5);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)  // dv[key][d] = sum_q P_d[q][key] dO[q][d]
                    umma_bf16(tmem + C_DV, mn_desc(s_p, kk), mn_desc(dO, kk), idesc_t, kk > 0);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)  // dk[key][d] = sum_q dS[q][key] q[q][d]
                    umma_bf16(tmem + C_DK, mn_desc(s_s, kk), mn_desc(q, kk), idesc_t, kk > 0);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)  // dq[q][d] = sum_key dS[q][key] k[key][d]
                    umma_bf16(tmem + C_DQ, k_desc128(s_s, kk), mn_desc(k, kk), idesc_q, kk > 0);
                umma_commit(o_ready);
                umma_commit(empty(b));
                BF_STAMP(1, 6);
            }
        }
    } else if (warp >= kRowWarps + 2) {
        // ===================== store warps: the staged gradient tiles (where P_d / dS were) -> global memory; on the way
        // (optional) their column sums = the bias gradients of the q / k / v projections: warp wb takes columns
        // 32 wb .. +31 of the three tiles (lane: word l % 16, rows r and r + 4 by l / 16 -- conflict-free under the
        // swizzle), adds them to this block's own partial row with fire-and-forget adds (one thread per address, adds in
        // program order: deterministic); a second small pass adds the rows in block order.
        const int wb = warp - (kRowWarps + 2);
        // lane: 16 B chunk 4 wb + l % 4 (8 columns) of row 8 it + rsel; rows r and r + 4 share a quarter warp (their
        // swizzled chunks fall into disjoint bank groups)
        const int cl = lane & 3, rsel = ((lane >> 3) & 3) + 4 * ((lane >> 2) & 1);
        const uint8_t* const tiles_gen = sm.gen + 2 * BWD_BUF;  // s_p as a generic pointer
        const int64_t HD = (int64_t)p.H * DD;
        for (int i = 0; i < n_mine; ++i) {
            const int pair = blockIdx.x + i * gridDim.x, bb = pair / p.H, h = pair - bb * p.H;
            mbar_wait(staged, i & 1);
            if (wb == 0 && lane == 0) {
                tma_store_4d(&map_dv, s_p, 0, 0, h, bb);
                tma_store_4d(&map_dk, s_p + TILE, 0, 0, h, bb);
                tma_store_4d(&map_dq, s_s, 0, 0, h, bb);
                tma_store_commit();
            }
            if (p.col_partial) {
                const int smp = bb / p.seqs_per_sample;
#pragma unroll 1
                for (int g = 0; g < 3; ++g) {  // staged order: dv, dk, dq -> rows 2, 1, 0 of [3][S][heads * 64]
                    const uint8_t* const tile = tiles_gen + g * TILE;
                    float a[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
                    for (int it = 0; it < 16; ++it) {
                        const int r = 8 * it + rsel;
                        const uint4 v = *reinterpret_cast<const uint4*>(tile + r * 128 + (((4 * wb + cl) ^ (r & 7)) << 4));
                        a[0] += __uint_as_float(v.x << 16), a[1] += __uint_as_float(v.x & 0xffff0000u);
                        a[2] += __uint_as_float(v.y << 16), a[3] += __uint_as_float(v.y & 0xffff0000u);
                        a[4] += __uint_as_float(v.z << 16), a[5] += __uint_as_float(v.z & 0xffff0000u);
                        a[6] += __uint_as_float(v.w << 16), a[7] += __uint_as_float(v.w & 0xffff0000u);
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) {  // the 8 row classes of a column: lanes l, l ^ 4, l ^ 8, l ^ 16
                        a[e] += __shfl_xor_sync(0xffffffffu, a[e], 4);
                        a[e] += __shfl_xor_sync(0xffffffffu, a[e], 8);
                        a[e] += __shfl_xor_sync(0xffffffffu, a[e], 16);
                    }
                    if (lane < 4) {
                        float* const dst = p.col_partial + (((int64_t)blockIdx.x * 3 + (2 - g)) * p.S + smp) * HD + (int64_t)h * DD +
                                           32 * wb + 8 * cl;
#pragma unroll
                        for (int e = 0; e < 8; ++e) atomicAdd(dst + e, a[e]);
                    }
                }
            }
            if (wb == 0 && lane == 0) tma_store_wait_read<0>();
            __syncwarp();
            if (lane == 0) mbar_arrive(stage_free);  // both warps: P_d / dS of the next pair may be written
        }
        if (wb == 0 && lane == 0) tma_store_wait_all();
    } else if (warp < kRowWarps) {
        const int q = 32 * (warp & 3) + lane, cq = warp >> 2;
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        const float ik = p.thresh != 0u ? p.inv_keep : 1.0f;
        uint32_t rp[32], rt[32];  // P (true probabilities) and t = keep ? dP_d : 0 of the pair in flight
        float part = 0.0f, lse_n = 0.0f;
        uint32_t kw = 0xffffffffu, kw_n = 0xffffffffu;
        auto fetch = [&](int i) {  // lse and keep word of pair i (used one pass later)
            if (i < n_mine) {
                const uint32_t grow = (uint32_t)(blockIdx.x + i * gridDim.x) * TT + q;
                lse_n = __ldg(p.lse + grow);
                if (p.thresh != 0u) kw_n = __ldg(p.keep + (size_t)grow * 4 + cq);
            }
        };
        auto pass1 = [&](int i) {  // S, dP of pair i -> P, t, partial D over this quarter of the keys
            const float lse = lse_n;
            kw = kw_n;
            mbar_wait(s_ready, i & 1);
            tc_fence_after();
            tmem_ld_32x32(lane_addr + C_S + 32 * cq, rp);
            tmem_ld_32x32(lane_addr + C_DP + 32 * cq, rt);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_free);
            fetch(i + 1);
            part = 0.0f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float pj = bf_ex2_approx(fmaf(__uint_as_float(rp[j]), p.scale_log2e, -lse));
                const float tj = (kw & (1u << j)) ? __uint_as_float(rt[j]) : 0.0f;
                part = fmaf(pj, tj, part);
                rp[j] = __float_as_uint(pj), rt[j] = __float_as_uint(tj);
            }
        };
        fetch(0);
        if (n_mine > 0) pass1(0);
        for (int i = 0; i < n_mine; ++i) {
            if (threadIdx.x == 0) BF_STAMP(0, 0);
            float* const xi = xch + (i & 1) * 4 * TT;  // by parity: this barrier is the only one of the pair
            xi[cq * TT + q] = part * ik;
            named_bar_sync<1, kRowThreads>();
            if (threadIdx.x == 0) BF_STAMP(0, 1);
            const float Dq = (xi[q] + xi[TT + q]) + (xi[2 * TT + q] + xi[3 * TT + q]);  // = rowsum(dO o O)
            // P_d = keep ? P / (1-p) : 0 and dS = P (t / (1-p) - D) scale, packed in place (rp <- P_d, rt <- dS) ...
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int j0 = 2 * e, j1 = j0 + 1;
                const float pa = __uint_as_float(rp[j0]), pb = __uint_as_float(rp[j1]);
                const uint32_t ds = pack_bf16(pa * p.scale * fmaf(__uint_as_float(rt[j0]), ik, -Dq),
                                              pb * p.scale * fmaf(__uint_as_float(rt[j1]), ik, -Dq));
                rp[e] = pack_bf16((kw & (1u << j0)) ? pa * ik : 0.0f, (kw & (1u << j1)) ? pb * ik : 0.0f);
                rt[e] = ds;
            }
            // ... while the gradient tiles of the previous pair leave the shared memory they were staged in; then row q of
            // key sub-tile cq / 2
            mbar_wait(stage_free, (i & 1) ^ 1u);
            {
                const uint32_t o0 = (cq >> 1) * TILE;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t o = o0 + chunk_off(q, 4 * (cq & 1) + c);
                    sts128(s_p + o, rp[4 * c], rp[4 * c + 1], rp[4 * c + 2], rp[4 * c + 3]);
                    sts128(s_s + o, rt[4 * c], rt[4 * c + 1], rt[4 * c + 2], rt[4 * c + 3]);
                }
            }
            fence_proxy_async();
            mbar_arrive(p_ready);
            if (threadIdx.x == 0) BF_STAMP(0, 2);
            if (i + 1 < n_mine) pass1(i + 1);  // under the dv / dk / dq MMAs of pair i
            if (threadIdx.x == 0) BF_STAMP(0, 3);
            // ---- gradients: row q (dq) / key q (dv, dk), columns 16 cq .. +15, staged where P_d / dS were (their MMAs
            //      are complete); one at a time: P, t of the next pair are live in registers
            mbar_wait(o_ready, i & 1);
            if (threadIdx.x == 0) BF_STAMP(0, 4);
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                uint32_t ro[16];
                tmem_ld_32x16(lane_addr + (g == 0 ? C_DV : g == 1 ? C_DK : C_DQ) + 16 * cq, ro);
                tmem_ld_wait();
                const uint32_t dst = g == 0 ? s_p : g == 1 ? s_p + TILE : s_s;
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    sts128(dst + chunk_off(q, 2 * cq + c), pack_bf16(__uint_as_float(ro[8 * c]), __uint_as_float(ro[8 * c + 1])),
                           pack_bf16(__uint_as_float(ro[8 * c + 2]), __uint_as_float(ro[8 * c + 3])),
                           pack_bf16(__uint_as_float(ro[8 * c + 4]), __uint_as_float(ro[8 * c + 5])),
                           pack_bf16(__uint_as_float(ro[8 * c + 6]), __uint_as_float(ro[8 * c + 7])));
            }
            tc_fence_before();
            mbar_arrive(o_free);
            fence_proxy_async();
            mbar_arrive(staged);
            if (threadIdx.x == 0) BF_STAMP(0, 5);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kRowWarps + 1) {
        tc_fence_after();
        tmem_dealloc(tmem, TMEM_COLS);
    }
}

// [B][128][heads][64]-indexed view (any batch / head / token strides, unit inner stride) as a 4-D tensor map:
// box = one (sequence, head) tile of 128 rows x 128 B, 128B swizzle
static int encode_pair_map(CUtensorMap* m, const void* base, int64_t B, int64_t H, int64_t sb, int64_t sh, int64_t st) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        bf_set_error("cuTensorMapEncodeTiled entry point not available");
        return BF_ERR_DRIVER;
    }
    const cuuint64_t dims[4] = {(cuuint64_t)DD, (cuuint64_t)TT, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)st * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
    const cuuint32_t box[4] = {(cuuint32_t)DD, (cuuint32_t)TT, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        bf_set_error("cuTensorMapEncodeTiled (attention tile) failed with CUresult " + std::to_string((int)r));
        return BF_ERR_DRIVER;
    }
    return 0;
}

// out[i] = sum over the per-block rows of the partial column sums, in block order (deterministic)
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                            int64_t n, int blocks) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float acc = 0.0f;
    for (int c = 0; c < blocks; ++c) acc += partial[(int64_t)c * n + i];
    out[i] = acc;
}

static void fill(Params& p, int64_t B, int64_t H, float scale, float p_drop, uint64_t seed, uint32_t step, uint32_t site,
                 float* lse, uint32_t* keep) {
    p.lse = lse, p.keep = keep;
    p.H = (int)H, p.total = (int)(B * H);
    p.scale = scale, p.scale_log2e = scale * 1.4426950408889634f;
    p.thresh = 0u, p.inv_keep = 1.0f;
    if (p_drop > 0.0f) {
        uint32_t t = (uint32_t)lrintf(p_drop * 65536.0f);
        if (t > 65535u) t = 65535u;
        p.thresh = t;
        p.inv_keep = 65536.0f / (65536.0f - (float)t);
    }
    p.k0 = (uint32_t)(seed & 0xffffffffu), p.k1 = (uint32_t)(seed >> 32), p.step = step, p.site = site;
    p.step_ptr = bf_step_counter();
}

}  // namespace attn_tc

// called by bf_attention_fwd / bf_attention_bwd (bf_attention.cu) after their argument checks, for T == 128
int bf_attention_tc_fwd(const void* q, const void* k, const void* v, const int64_t* strides, int64_t B, int64_t H, float scale,
                        float p_drop, uint64_t seed, uint32_t step, uint32_t site, void* out, float* lse, uint32_t* keep,
                        cudaStream_t stream) {
    using namespace attn_tc;
    CUtensorMap mq, mk, mv;
    int rc;
    if ((rc = encode_pair_map(&mq, q, B, H, strides[0], strides[1], strides[2]))) return rc;
    if ((rc = encode_pair_map(&mk, k, B, H, strides[3], strides[4], strides[5]))) return rc;
    if ((rc = encode_pair_map(&mv, v, B, H, strides[6], strides[7], strides[8]))) return rc;
    Params p{};
    fill(p, B, H, scale, p_drop, seed, step, site, lse, keep);
    p.out = (__nv_bfloat16*)out;
    BF_CUDA_OK(cudaFuncSetAttribute(fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    const int grid = p.total < bf_num_sms() ? p.total : bf_num_sms();
    fwd_kernel<<<grid, kFwdThreads, FWD_SMEM, stream>>>(mq, mk, mv, p);
    BF_LAUNCH_OK();
    return 0;
}

int bf_attention_tc_bwd(const void* dout, const void* q, const void* k, const void* v, const int64_t* strides, const float* lse,
                        const uint32_t* keep, int64_t B, int64_t H, float scale, float p_drop, uint64_t seed, uint32_t step,
                        uint32_t site, void* dq, void* dk, void* dv, float* dbias, void* workspace, int64_t S,
                        cudaStream_t stream) {
    using namespace attn_tc;
    CUtensorMap mq, mk, mv, mdo, mdq, mdk, mdv;
    const int64_t osb = (int64_t)TT * H * DD, osh = DD, ost = H * DD;
    int rc;
    if ((rc = encode_pair_map(&mq, q, B, H, strides[0], strides[1], strides[2]))) return rc;
    if ((rc = encode_pair_map(&mk, k, B, H, strides[3], strides[4], strides[5]))) return rc;
    if ((rc = encode_pair_map(&mv, v, B, H, strides[6], strides[7], strides[8]))) return rc;
    if ((rc = encode_pair_map(&mdo, dout, B, H, osb, osh, ost))) return rc;
    if ((rc = encode_pair_map(&mdq, dq, B, H, osb, osh, ost))) return rc;
    if ((rc = encode_pair_map(&mdk, dk, B, H, osb, osh, ost))) return rc;
    if ((rc = encode_pair_map(&mdv, dv, B, H, osb, osh, ost))) return rc;
    Params p{};
    fill(p, B, H, scale, p_drop, seed, step, site, const_cast<float*>(lse), const_cast<uint32_t*>(keep));
    BF_CUDA_OK(cudaFuncSetAttribute(bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    const int grid = p.total < bf_num_sms() ? p.total : bf_num_sms();
    const int64_t n_bias = 3 * S * H * DD;
    if (dbias) {
        p.col_partial = static_cast<float*>(workspace), p.S = (int)S, p.seqs_per_sample = (int)(B / S);
        BF_CUDA_OK(cudaMemsetAsync(workspace, 0, (size_t)grid * n_bias * 4, stream));
    }
    bwd_kernel<<<grid, kBwdThreads, BWD_SMEM, stream>>>(mq, mk, mv, mdo, mdq, mdk, mdv, p);
    BF_LAUNCH_OK();
    if (dbias) {
        colsum_reduce_kernel<<<(unsigned)((n_bias + 255) / 256), 256, 0, stream>>>(p.col_partial, dbias, n_bias, grid);
        BF_LAUNCH_OK();
    }
    return 0;
}

int64_t bf_attention_tc_bias_workspace_bytes(int64_t B, int64_t H, int64_t S) {
    const int64_t total = B * H, grid = total < bf_num_sms() ? total : bf_num_sms();
    return grid * 3 * S * H * attn_tc::DD * 4;
}
