// S-sample Linear contractions on 5th-gen tensor cores (sm_100a):
// TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring -> tcgen05.mma
// (bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld epilogue, warp-specialised and
// persistent (one CTA per SM).
//
// Replaces F.linear (bayeformers/nn/layers/linear.py:104) and its autograd for
// all S Monte-Carlo samples at once.  The three contractions share one kernel:
//
//     D[s][i][j] = sum_r A_s(i, r) * B_s(j, r)
//
//   fwd   : D = y  [M,N]   A = x  (K-major)   B = w  (K-major)   r = K
//   dgrad : D = dx [M,K]   A = gy (K-major)   B = w  (MN-major)  r = N
//   wgrad : D = dW [N,K]   A = gy (MN-major)  B = x  (MN-major)  r = M
//
// "K-major" = the reduction index is the contiguous one in memory; "MN-major"
// = the output index is contiguous (no transposes are ever materialised).
//
// wgrad's epilogue is the variational backward: the fp32 accumulator tile of
// sample s is combined with a regenerated Philox eps tile and folded into
// grad_mu / grad_rho (softplus chain + optional KL gradient), so the S weight
// gradients never exist in HBM.  Work items that update the same output tile
// (different samples / reduction splits) take turns in a fixed order, which
// keeps the result run-to-run deterministic without float atomics.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "bf_common.cuh"

namespace tc {

constexpr int BLOCK_M = 128;   // UMMA M (TMEM lanes)
constexpr int BLOCK_N = 256;   // UMMA N (TMEM columns per accumulator)
constexpr int BLOCK_K = 64;    // reduction elements per pipeline stage (= one 128 B swizzle row of bf16)
constexpr int UMMA_K = 16;     // bf16 reduction depth of one tcgen05.mma
constexpr int kStages = 4;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;  // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int ATOM_MN = 64;                           // MN-major: 64 contiguous bf16 = 128 B swizzle atom
constexpr int ATOM_BYTES = ATOM_MN * BLOCK_K * 2;     // one MN-major TMA box: 64 rows x 128 B = 8 KiB
constexpr int EPI_COLS = 32;                          // accumulator columns per epilogue chunk
constexpr int EPI_STRIDE = 36;                        // floats per staged row (16 B aligned, conflict-free)
constexpr int EPI_WARPS = 4;
constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_STRIDE * 4;
constexpr int kThreads = 32 * (2 + EPI_WARPS);        // warp0 TMA, warp1 MMA, warps 2..5 epilogue
constexpr int TMEM_COLS = 2 * BLOCK_N;                // double-buffered accumulator = all 512 columns
constexpr int SMEM_BYTES = 1024 /*align slack*/ + kStages * STAGE_BYTES + EPI_BYTES + 256;

enum { EPI_STORE = 0, EPI_VARGRAD = 1 };

struct Params {
    int64_t S, I, J, R;  // batch, rows of D, cols of D, reduction length
    int i_tiles, j_tiles, k_steps, splits;
    // EPI_STORE
    void* out;
    int out_dtype;
    const float* bias;  // [S][J] or null
    // EPI_VARGRAD
    const float* mu;
    const float* rho;
    const float* prior_mu;
    const float* prior_rho;
    const float* g_logq;
    const float* g_logp;
    const float* eps_in;
    float* grad_mu;
    float* grad_rho;
    int* turn;  // [i_tiles*j_tiles], zero-initialised, self-resetting
    int accumulate;
    int prior_kind;
    uint32_t k0, k1, step, tensor_id;
    float prior_const_ipv;  // 1/sigma_p^2 for a Gaussian prior with constant sigma (prior_rho == NULL)
    BfMixture mix;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (lane = accumulator row)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ descriptors
// UMMA shared-memory matrix descriptor, 128B swizzle (sm_100 format, version 1):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor, kind::f16: D=f32 (bits[4,6)=1), A=B=bf16 (bits[7,10)=[10,13)=1),
// a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

struct Item {
    int s, i_blk, j_blk, k_begin, k_end, tile, turn, turns;
};

template <int EPI>
__device__ __forceinline__ Item decode_item(const Params& p, int64_t L) {
    Item it;
    if (EPI == EPI_STORE) {
        it.j_blk = (int)(L % p.j_tiles);
        const int64_t q = L / p.j_tiles;
        it.i_blk = (int)(q % p.i_tiles);
        it.s = (int)(q / p.i_tiles);
        it.k_begin = 0, it.k_end = p.k_steps;
        it.tile = 0, it.turn = 0, it.turns = 1;
    } else {
        const int turns = (int)p.S * p.splits;
        it.turns = turns;
        it.turn = (int)(L % turns);
        it.tile = (int)(L / turns);
        it.s = it.turn / p.splits;
        const int split = it.turn % p.splits;
        it.j_blk = it.tile % p.j_tiles;
        it.i_blk = it.tile / p.j_tiles;
        it.k_begin = (int)(((int64_t)p.k_steps * split) / p.splits);
        it.k_end = (int)(((int64_t)p.k_steps * (split + 1)) / p.splits);
    }
    return it;
}

__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// ------------------------------------------------------------------ the kernel
template <bool A_MN, bool B_MN, int EPI, bool KL>
__global__ void __launch_bounds__(kThreads, 1)
    bayes_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B swizzle needs 1024 B alignment
    uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    float* const epi_stage = reinterpret_cast<float*>(smem_gen + kStages * STAGE_BYTES);
    const uint32_t bar_base = smem_base + kStages * STAGE_BYTES + EPI_BYTES;
    // barriers: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then the TMEM base address slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    volatile uint32_t* const tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * STAGE_BYTES + EPI_BYTES + 8 * (2 * kStages + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), EPI_WARPS);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    const int64_t n_items = (EPI == EPI_STORE) ? p.S * p.i_tiles * p.j_tiles
                                               : (int64_t)p.i_tiles * p.j_tiles * p.S * p.splits;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t L = blockIdx.x; L < n_items; L += gridDim.x) {
                const Item it = decode_item<EPI>(p, L);
                const int i0 = it.i_blk * BLOCK_M, j0 = it.j_blk * BLOCK_N;
                for (int ks = it.k_begin; ks < it.k_end; ++ks) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
                    const uint32_t b_dst = a_dst + A_BYTES;
                    mbar_expect_tx(full_bar(stage), STAGE_BYTES);
                    const int r0 = ks * BLOCK_K;
                    if (A_MN) {
#pragma unroll
                        for (int a = 0; a < BLOCK_M / ATOM_MN; ++a)
                            tma_load_3d(a_dst + a * ATOM_BYTES, &map_a, full_bar(stage), i0 + a * ATOM_MN, r0, it.s);
                    } else {
                        tma_load_3d(a_dst, &map_a, full_bar(stage), r0, i0, it.s);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int a = 0; a < BLOCK_N / ATOM_MN; ++a)
                            tma_load_3d(b_dst + a * ATOM_BYTES, &map_b, full_bar(stage), j0 + a * ATOM_MN, r0, it.s);
                    } else {
                        tma_load_3d(b_dst, &map_b, full_bar(stage), r0, j0, it.s);
                    }
                    if (++stage == kStages) stage = 0, phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(A_MN, B_MN);
            // K-major : rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused.  +32 B per UMMA_K step.
            // MN-major: 64-wide atoms ATOM_BYTES apart (LBO), 8 reduction rows per 1024 B group (SBO);
            //           +16 rows = 2048 B per UMMA_K step.
            constexpr uint32_t a_lbo = A_MN ? ATOM_BYTES : 16, a_sbo = 1024, a_kadv = A_MN ? 2048 : 32;
            constexpr uint32_t b_lbo = B_MN ? ATOM_BYTES : 16, b_sbo = 1024, b_kadv = B_MN ? 2048 : 32;
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int64_t L = blockIdx.x; L < n_items; L += gridDim.x, ++iter) {
                const Item it = decode_item<EPI>(p, L);
                const int acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int ks = it.k_begin; ks < it.k_end; ++ks) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_src = smem_base + stage * STAGE_BYTES;
                    const uint32_t b_src = a_src + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t ad = make_smem_desc(a_src + k * a_kadv, a_lbo, a_sbo);
                        const uint64_t bd = make_smem_desc(b_src + k * b_kadv, b_lbo, b_sbo);
                        umma_bf16(d_tmem, ad, bd, idesc, (ks > it.k_begin || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs retire
                    if (++stage == kStages) stage = 0, phase ^= 1u;
                }
                umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may touch
        float* const stage_w = epi_stage + q * 32 * EPI_STRIDE;
        int iter = 0;
        for (int64_t L = blockIdx.x; L < n_items; L += gridDim.x, ++iter) {
            const Item it = decode_item<EPI>(p, L);
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int64_t i_base = (int64_t)it.i_blk * BLOCK_M + q * 32;
            const int64_t j_base = (int64_t)it.j_blk * BLOCK_N;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            if (EPI == EPI_VARGRAD) {
                // fixed-order turn taking between the items that update this output tile
                if (lane == 0) {
                    const volatile int* t = p.turn + it.tile;
                    while (*t != it.turn) __nanosleep(64);
                    __threadfence();
                }
                __syncwarp();
            }
            float glq = 0.0f, glp = 0.0f;
            if (EPI == EPI_VARGRAD && KL && it.k_begin == 0) {
                glq = p.g_logq ? __ldg(p.g_logq + it.s) : 0.0f;
                glp = p.g_logp ? __ldg(p.g_logp + it.s) : 0.0f;
            }
            const bool first_turn = (it.turn == 0);
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / EPI_COLS; ++c) {
                const int64_t jc = j_base + c * EPI_COLS;
                if (jc >= p.J) break;  // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c * EPI_COLS), r);
                // stage: lane = row, 8 x 16 B per row (conflict-free with the 36-float row stride)
#pragma unroll
                for (int v = 0; v < 8; ++v)
                    *reinterpret_cast<uint4*>(stage_w + lane * EPI_STRIDE + 4 * v) =
                        make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
                __syncwarp();
                // coalesced phase: 8 lanes x float4 cover one 32-column row segment, 4 rows per pass
                const int cj = 4 * (lane & 7);
                const int64_t j = jc + cj;
#pragma unroll 2
                for (int pass = 0; pass < 8; ++pass) {
                    const int rr = pass * 4 + (lane >> 3);
                    const int64_t i = i_base + rr;
                    if (i >= p.I || j >= p.J) continue;
                    const float4 a = ld_f4(stage_w + rr * EPI_STRIDE + cj);
                    float v[4] = {a.x, a.y, a.z, a.w};
                    const bool full = (j + 4 <= p.J);
                    if (EPI == EPI_STORE) {
                        if (p.bias != nullptr) {
                            const float* bp = p.bias + (int64_t)it.s * p.J + j;
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (j + e < p.J) v[e] += __ldg(bp + e);
                        }
                        const int64_t off = ((int64_t)it.s * p.I + i) * p.J + j;
                        if (p.out_dtype == BF_BF16) {
                            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
                            if (full && ((p.J & 3) == 0)) {
                                const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]);
                                const __nv_bfloat162 hi = __floats2bfloat162_rn(v[2], v[3]);
                                uint2 u;
                                u.x = *reinterpret_cast<const uint32_t*>(&lo);
                                u.y = *reinterpret_cast<const uint32_t*>(&hi);
                                *reinterpret_cast<uint2*>(o) = u;
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (j + e < p.J) o[e] = __float2bfloat16_rn(v[e]);
                            }
                        } else {
                            float* o = reinterpret_cast<float*>(p.out) + off;
                            if (full && ((p.J & 3) == 0)) {
                                *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (j + e < p.J) o[e] = v[e];
                            }
                        }
                    } else {
                        // ---- variational backward on a 4-element quad of W[i][j..j+3] (J % 4 == 0 guaranteed)
                        const int64_t flat = i * p.J + j;
                        const float4 rho4 = __ldg(reinterpret_cast<const float4*>(p.rho + flat));
                        const float rho[4] = {rho4.x, rho4.y, rho4.z, rho4.w};
                        float e4[4];
                        if (p.eps_in != nullptr) {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(
                                p.eps_in + (int64_t)it.s * p.I * p.J + flat));
                            e4[0] = t.x, e4[1] = t.y, e4[2] = t.z, e4[3] = t.w;
                        } else {
                            const float4 t = bf_eps_quad((uint32_t)(flat >> 2), (uint32_t)it.s, p.tensor_id, p.step,
                                                         p.k0, p.k1);
                            e4[0] = t.x, e4[1] = t.y, e4[2] = t.z, e4[3] = t.w;
                        }
                        float gm[4], gr[4];
                        if (KL && it.k_begin == 0) {
                            const float4 mu4 = __ldg(reinterpret_cast<const float4*>(p.mu + flat));
                            const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w};
                            float pm[4] = {0, 0, 0, 0}, ipv[4] = {0, 0, 0, 0};
                            if (p.prior_kind == BF_PRIOR_GAUSSIAN) {
                                const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.prior_mu + flat));
                                pm[0] = a4.x, pm[1] = a4.y, pm[2] = a4.z, pm[3] = a4.w;
                                if (p.prior_rho != nullptr) {
                                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.prior_rho + flat));
                                    const float pr[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float sp = bf_softplus(pr[e]);
                                        ipv[e] = 1.0f / (sp * sp);
                                    }
                                } else {
                                    ipv[0] = ipv[1] = ipv[2] = ipv[3] = p.prior_const_ipv;
                                }
                            }
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float sg = bf_softplus(rho[e]);
                                const float w = __fadd_rn(mu[e], __fmul_rn(e4[e], sg));
                                float dp = 0.0f;
                                if (p.prior_kind == BF_PRIOR_MIXTURE) dp = bf_mixture_dlogp(w, p.mix);
                                if (p.prior_kind == BF_PRIOR_GAUSSIAN) dp = -(w - pm[e]) * ipv[e];
                                const float g = v[e] + glp * dp;
                                gm[e] = g;
                                gr[e] = (g * e4[e] - glq / sg) * bf_softplus_grad(rho[e]);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                gm[e] = v[e];
                                gr[e] = v[e] * e4[e] * bf_softplus_grad(rho[e]);
                            }
                        }
                        const bool add = !first_turn || p.accumulate;
                        float4* dr = reinterpret_cast<float4*>(p.grad_rho + flat);
                        float4 o = make_float4(gr[0], gr[1], gr[2], gr[3]);
                        if (add) {
                            const float4 old = __ldcg(dr);
                            o.x += old.x, o.y += old.y, o.z += old.z, o.w += old.w;
                        }
                        __stcg(dr, o);
                        if (p.grad_mu != nullptr) {
                            float4* dm = reinterpret_cast<float4*>(p.grad_mu + flat);
                            float4 om = make_float4(gm[0], gm[1], gm[2], gm[3]);
                            if (add) {
                                const float4 old = __ldcg(dm);
                                om.x += old.x, om.y += old.y, om.z += old.z, om.w += old.w;
                            }
                            __stcg(dm, om);
                        }
                    }
                }
                __syncwarp();
            }
            // accumulator drained: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (EPI == EPI_VARGRAD) {
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
                if (warp == 2 && lane == 0) {
                    const int next = (it.turn + 1 == it.turns) ? 0 : it.turn + 1;
                    __threadfence();
                    atomicExch(p.turn + it.tile, next);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

// bf16 tensor [S][rows][cols] (cols contiguous); box = {64 cols, box_rows, 1}, 128B swizzle, zero OOB fill
static int encode_map(CUtensorMap* m, const void* base, int64_t S, int64_t rows, int64_t cols, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        bf_set_error("cuTensorMapEncodeTiled entry point not available");
        return BF_ERR_DRIVER;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)S};
    const cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * (cuuint64_t)cols * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        bf_set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return BF_ERR_DRIVER;
    }
    return 0;
}

template <bool A_MN, bool B_MN, int EPI, bool KL>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const Params& p, int64_t n_items, cudaStream_t st) {
    auto kern = bayes_gemm_kernel<A_MN, B_MN, EPI, KL>;
    BF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int64_t sms = bf_num_sms();
    const int grid = (int)(n_items < sms ? n_items : sms);
    kern<<<grid, kThreads, SMEM_BYTES, st>>>(ma, mb, p);
    BF_LAUNCH_OK();
    return 0;
}

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace tc

// y[s] = x[s] w[s]^T + bias[s]
int bf_linear_fwd_bf16(const void* x, const void* w, const float* bias, void* y, int64_t S, int64_t M, int64_t N,
                       int64_t K, int32_t y_dtype, cudaStream_t st) {
    using namespace tc;
    BF_CHECK_ARG(K % 8 == 0, "bf16 path needs K % 8 == 0 (16 B TMA row stride)");
    CUtensorMap ma, mb;
    int rc;
    if ((rc = encode_map(&ma, x, S, M, K, BLOCK_M))) return rc;
    if ((rc = encode_map(&mb, w, S, N, K, BLOCK_N))) return rc;
    Params p{};
    p.S = S, p.I = M, p.J = N, p.R = K;
    p.i_tiles = cdiv(M, BLOCK_M), p.j_tiles = cdiv(N, BLOCK_N), p.k_steps = cdiv(K, BLOCK_K), p.splits = 1;
    p.out = y, p.out_dtype = y_dtype, p.bias = bias;
    return launch<false, false, EPI_STORE, false>(ma, mb, p, S * p.i_tiles * p.j_tiles, st);
}

// dx[s] = gy[s] w[s]
int bf_linear_dgrad_bf16(const void* gy, const void* w, void* dx, int64_t S, int64_t M, int64_t N, int64_t K,
                         int32_t dx_dtype, cudaStream_t st) {
    using namespace tc;
    BF_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "bf16 path needs K % 8 == 0 and N % 8 == 0");
    CUtensorMap ma, mb;
    int rc;
    if ((rc = encode_map(&ma, gy, S, M, N, BLOCK_M))) return rc;      // A: K-major over r = N
    if ((rc = encode_map(&mb, w, S, N, K, BLOCK_K))) return rc;       // B: MN-major, rows = r = N, cols = K
    Params p{};
    p.S = S, p.I = M, p.J = K, p.R = N;
    p.i_tiles = cdiv(M, BLOCK_M), p.j_tiles = cdiv(K, BLOCK_N), p.k_steps = cdiv(N, BLOCK_K), p.splits = 1;
    p.out = dx, p.out_dtype = dx_dtype, p.bias = nullptr;
    return launch<false, true, EPI_STORE, false>(ma, mb, p, S * p.i_tiles * p.j_tiles, st);
}

// dw[s] = gy[s]^T x[s]   (plain store, fp32 [S,N,K])
int bf_linear_wgrad_bf16(const void* gy, const void* x, float* dw, int64_t S, int64_t M, int64_t N, int64_t K,
                         cudaStream_t st) {
    using namespace tc;
    BF_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "bf16 path needs K % 8 == 0 and N % 8 == 0");
    CUtensorMap ma, mb;
    int rc;
    if ((rc = encode_map(&ma, gy, S, M, N, BLOCK_K))) return rc;      // A: MN-major, rows = r = M, cols = N
    if ((rc = encode_map(&mb, x, S, M, K, BLOCK_K))) return rc;       // B: MN-major, rows = r = M, cols = K
    Params p{};
    p.S = S, p.I = N, p.J = K, p.R = M;
    p.i_tiles = cdiv(N, BLOCK_M), p.j_tiles = cdiv(K, BLOCK_N), p.k_steps = cdiv(M, BLOCK_K), p.splits = 1;
    p.out = dw, p.out_dtype = BF_F32, p.bias = nullptr;
    return launch<true, true, EPI_STORE, false>(ma, mb, p, S * p.i_tiles * p.j_tiles, st);
}

int64_t bf_wgrad_fused_workspace_ints(int64_t N, int64_t K) {
    return (int64_t)tc::cdiv(N, tc::BLOCK_M) * tc::cdiv(K, tc::BLOCK_N);
}

int bf_linear_wgrad_fused_bf16(const void* gy, const void* x, int64_t S, int64_t M, int64_t N, int64_t K,
                               const float* mu, const float* rho, int32_t prior_kind, const float* prior_mu,
                               const float* prior_rho, float pi, float sigma1, float sigma2, const float* g_logq,
                               const float* g_logp, uint64_t seed, uint32_t step, uint32_t tensor_id,
                               const float* eps_in, float* grad_mu, float* grad_rho, int32_t accumulate, int* turn_ws,
                               cudaStream_t st) {
    using namespace tc;
    BF_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "bf16 path needs K % 8 == 0 and N % 8 == 0");
    BF_CHECK_ARG(turn_ws != nullptr, "turn workspace missing");
    CUtensorMap ma, mb;
    int rc;
    if ((rc = encode_map(&ma, gy, S, M, N, BLOCK_K))) return rc;
    if ((rc = encode_map(&mb, x, S, M, K, BLOCK_K))) return rc;
    Params p{};
    p.S = S, p.I = N, p.J = K, p.R = M;
    p.i_tiles = cdiv(N, BLOCK_M), p.j_tiles = cdiv(K, BLOCK_N), p.k_steps = cdiv(M, BLOCK_K);
    // split the reduction until the persistent grid is reasonably full
    const int64_t tiles = (int64_t)p.i_tiles * p.j_tiles;
    int splits = 1;
    const int64_t sms = bf_num_sms();
    while (tiles * S * splits * 2 <= sms && p.k_steps / (splits * 2) >= 8) splits *= 2;
    p.splits = splits;
    p.mu = mu, p.rho = rho, p.prior_mu = prior_mu, p.prior_rho = prior_rho;
    p.g_logq = g_logq, p.g_logp = g_logp, p.eps_in = eps_in;
    p.grad_mu = grad_mu, p.grad_rho = grad_rho, p.turn = turn_ws, p.accumulate = accumulate;
    p.prior_kind = prior_kind;
    p.k0 = (uint32_t)(seed & 0xffffffffu), p.k1 = (uint32_t)(seed >> 32), p.step = step, p.tensor_id = tensor_id;
    p.mix = bf_make_mixture(pi, sigma1, sigma2);
    p.prior_const_ipv = 1.0f / (sigma1 * sigma1);
    const bool kl = (g_logq != nullptr) || (g_logp != nullptr);
    const int64_t n_items = tiles * S * splits;
    if (kl) return launch<true, true, EPI_VARGRAD, true>(ma, mb, p, n_items, st);
    return launch<true, true, EPI_VARGRAD, false>(ma, mb, p, n_items, st);
}
