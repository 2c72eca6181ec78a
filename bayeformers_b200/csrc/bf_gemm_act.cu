// S-sample Linear forward with the bias + GELU epilogue fused (CTA-pair tcgen05 kernel), and the matching
// fused backward elementwise pass.
//
// Extension of bnn.Linear (`activation="gelu"`): y = gelu(x w^T + b).  In the reference the activation is a
// separate module after the layer (F.linear, bayeformers/nn/layers/linear.py:104, followed by the host model's
// GELU); fusing it into the producing kernel removes one full read + write of the [S*M, N] activation in the
// forward pass and, in the backward pass, merges GELU' with the bias-gradient column sums (one pass over gy).
//
//   forward : z = x w^T + b (bf16, kept for backward),  y = gelu(z) (bf16)     -- both TMA-stored from the epilogue
//   backward: gz = gy * gelu'(z),  db[s][n] = sum_m gz[s][m][n]                -- bf_gelu_bwd_bias_grad
//
// GELU is the exact (erf) form HF BERT uses.  erf comes from Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7, far below
// bf16 resolution), evaluated on packed fp32 pairs (FFMA2).  Measured at the BERT-base FFN shape (S=4, M=32768,
// 3072x768): plain forward 448 us, this kernel 745 us (598 us with the GELU arithmetic removed, i.e. the second
// 805 MB result write costs 150 us and the epilogue arithmetic another 150 us: with 12 k-steps per tile the epilogue,
// not the MMA stream, is the critical path), against 448 + ~440 us for the separate GELU pass it replaces.
//
// Kernel structure = bf_gemm_tc2.cu (cluster of 2, cta_group::2, UMMA 256x256x16, leader issues, multicast commits)
// with 8 epilogue warps in two groups (group g takes the 64-column boxes b with b % 2 == g), 5 stages, and two
// staging buffers (z, y) per group.
#include "bf_tc.cuh"

namespace act {
using namespace tc;

// phase timeline of cluster 0's leader (scripts/gelu_trace.cu builds this file with -DBF_GELU_TRACE); compiled out otherwise
#ifdef BF_GELU_TRACE
__device__ unsigned long long g_gelu_trace[3 * 16 * 16];
#define BF_GSTAMP(who, it, slot) \
    do { if (blockIdx.x == 0 && (it) < 16) g_gelu_trace[((who) * 16 + (it)) * 16 + (slot)] = clock64(); } while (0)
#else
#define BF_GSTAMP(who, it, slot) do { } while (0)
#endif

constexpr int BLOCK_M = 128, BLOCK_N = 256, LOAD_N = 128;
constexpr int kStages = 5;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2, B_BYTES = LOAD_N * BLOCK_K * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int BOX_COLS = 64;                 // bf16 columns per TMA-store box (128 B rows)
constexpr int BOX_BYTES = BLOCK_M * 128;     // 16 KiB
constexpr int BOXES = BLOCK_N / BOX_COLS;    // 4
constexpr int EPI_GROUPS = 2, EPI_WARPS = 4 * EPI_GROUPS;
constexpr int kThreads = 32 * (2 + EPI_WARPS);
constexpr int TMEM_COLS = 2 * BLOCK_N;
constexpr int OUT_BYTES = EPI_GROUPS * 2 * BOX_BYTES;  // (z, y) per group
constexpr int HB_COLS = 32, HB_BYTES = BLOCK_M * HB_COLS * 2;  // half box of the epilogue staging: 64 B rows, 8 KiB
static_assert(OUT_BYTES == EPI_GROUPS * 4 * HB_BYTES, "two (z, y) half-box sets per group");
constexpr int SMEM_BYTES = 1024 + kStages * STAGE_BYTES + OUT_BYTES + 256;

struct Params {
    int64_t S, I, J, R;
    int i_pairs, j_tiles, k_steps;
    const float* bias;  // [S][J], required
    // dgelu only: per-block partial column sums of the result, [gridDim.x][S][J] fp32 (null: not wanted).  Every block
    // adds the column sums of ITS tiles, in its fixed tile order, to its own row; a second pass adds the rows in order.
    float* col_partial;
};
struct Item {
    int s, i_pair, j_blk;
};
__device__ __forceinline__ Item decode_item(const Params& p, int64_t L) {
    Item it;
    it.j_blk = (int)(L % p.j_tiles);
    const int64_t q = L / p.j_tiles;
    it.i_pair = (int)(q % p.i_pairs);
    it.s = (int)(q / p.i_pairs);
    return it;
}

// erf(x), Abramowitz & Stegun 7.1.26, |abs err| <= 1.5e-7
__device__ __forceinline__ float erf_as(float x) {
    const float ax = fabsf(x);
    const float t = bf_rcp_approx(fmaf(0.3275911f, ax, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = bf_ex2_approx(-1.4426950408889634f * ax * ax);
    const float r = fmaf(-p * t, e, 1.0f);
    return copysignf(r, x);
}
__device__ __forceinline__ float gelu_erf(float z) { return 0.5f * z * (1.0f + erf_as(z * 0.70710678118654752f)); }
// the same GELU on two values at once with packed fp32 arithmetic (FFMA2 / FMUL2): the forward epilogue has to stay
// cheaper than the 12 k-steps of MMA of a 768-deep tile, and this halves its FMA issue count
__device__ __forceinline__ bf_f2 gelu_erf2(bf_f2 z2) {
    // forward only needs erf itself (no exponential to share with a density), so it uses Abramowitz & Stegun 7.1.28,
    //   erf(x) = 1 - (1 + a1 x + ... + a6 x^6)^-16,  |abs err| <= 3e-7,
    // one MUFU (rcp) per value instead of 7.1.26's two (rcp + ex2): the epilogue's MUFU traffic was 2/3 of a tile's MMA time
    float x0, x1, q0, q1, r0, r1;
    bf_unpack2(bf_mul2(z2, bf_splat2(0.70710678118654752f)), x0, x1);
    const bf_f2 ax2 = bf_pack2(fabsf(x0), fabsf(x1));
    bf_f2 p2 = bf_fma2(bf_splat2(0.0000430638f), ax2, bf_splat2(0.0002765672f));
    p2 = bf_fma2(p2, ax2, bf_splat2(0.0001520143f));
    p2 = bf_fma2(p2, ax2, bf_splat2(0.0092705272f));
    p2 = bf_fma2(p2, ax2, bf_splat2(0.0422820123f));
    p2 = bf_fma2(p2, ax2, bf_splat2(0.0705230784f));
    p2 = bf_fma2(p2, ax2, bf_splat2(1.0f));
    p2 = bf_mul2(p2, p2);
    p2 = bf_mul2(p2, p2);
    p2 = bf_mul2(p2, p2);
    bf_unpack2(bf_mul2(p2, p2), q0, q1);  // ^16; overflows to +inf for |x| > ~17, where rcp gives 0 and erf = 1
    r0 = 1.0f - bf_rcp_approx(q0), r1 = 1.0f - bf_rcp_approx(q1);  // |erf|
    const bf_f2 erf2 = bf_pack2(copysignf(r0, x0), copysignf(r1, x1));
    return bf_mul2(z2, bf_fma2(erf2, bf_splat2(0.5f), bf_splat2(0.5f)));
}
// The two fused epilogues run under the board's power cap, where what a tile costs is the ENERGY of its instructions,
// not their issue slots (halving the epilogue's stall cycles left the kernel time unchanged: the clock dropped instead).
// So the epilogues use the cheapest evaluation that stays far inside bf16's rounding (2^-9 = 3.9e-3 relative): odd
// polynomials in z on the clamped argument, FFMA2 only -- no MUFU, no |z| / copysign.
//   Phi(z) - 1/2 = z Q(z^2),  |z| <= 4: degree 15, |abs err| <= 2.3e-5 in fp32 (least-maximum fit, Lawson iteration;
//   scripts/fit_gelu_poly.py); beyond +-4 the clamped value is used: |Phi(z) - Phi(+-4)| <= 3.2e-5 (5.3e-5 in total).
__device__ __forceinline__ bf_f2 gelu_poly2(bf_f2 z2) {
    float z0, z1;
    bf_unpack2(z2, z0, z1);
    const bf_f2 c2 = bf_pack2(fminf(fmaxf(z0, -4.0f), 4.0f), fminf(fmaxf(z1, -4.0f), 4.0f));
    const bf_f2 t2 = bf_mul2(c2, c2);
    bf_f2 q2 = bf_fma2(bf_splat2(-1.5807585973e-09f), t2, bf_splat2(1.2170958112e-07f));
    q2 = bf_fma2(q2, t2, bf_splat2(-4.1008329283e-06f));
    q2 = bf_fma2(q2, t2, bf_splat2(8.0667023899e-05f));
    q2 = bf_fma2(q2, t2, bf_splat2(-1.0482022168e-03f));
    q2 = bf_fma2(q2, t2, bf_splat2(9.6648676795e-03f));
    q2 = bf_fma2(q2, t2, bf_splat2(-6.6175370767e-02f));
    q2 = bf_fma2(q2, t2, bf_splat2(3.9884750669e-01f));
    return bf_mul2(z2, bf_fma2(c2, q2, bf_splat2(0.5f)));  // z Phi(z)
}
//   gelu'(z) - 1/2 = Phi(z) - 1/2 + z phi(z) = z R(z^2),  |z| <= 4: degree 17, |abs err| <= 9.1e-5 in fp32; beyond +-4
//   the clamped value is used: |gelu'(z) - gelu'(+-4)| <= 5.4e-4 (gelu'(4) = 1.0005, gelu'(-4) = -0.0005).
__device__ __forceinline__ bf_f2 gelu_grad_poly2(bf_f2 z2) {
    float z0, z1;
    bf_unpack2(z2, z0, z1);
    const bf_f2 c2 = bf_pack2(fminf(fmaxf(z0, -4.0f), 4.0f), fminf(fmaxf(z1, -4.0f), 4.0f));
    const bf_f2 t2 = bf_mul2(c2, c2);
    bf_f2 r2 = bf_fma2(bf_splat2(9.795993143e-10f), t2, bf_splat2(-8.218767033e-08f));
    r2 = bf_fma2(r2, t2, bf_splat2(3.028339970e-06f));
    r2 = bf_fma2(r2, t2, bf_splat2(-6.495748858e-05f));
    r2 = bf_fma2(r2, t2, bf_splat2(9.073261046e-04f));
    r2 = bf_fma2(r2, t2, bf_splat2(-8.716319043e-03f));
    r2 = bf_fma2(r2, t2, bf_splat2(5.845609926e-02f));
    r2 = bf_fma2(r2, t2, bf_splat2(-2.648265343e-01f));
    r2 = bf_fma2(r2, t2, bf_splat2(7.976095509e-01f));
    return bf_fma2(c2, r2, bf_splat2(0.5f));
}
// d gelu / dz = Phi(z) + z * phi(z) on two values; phi shares its exponential with erf: exp(-z^2/2) = exp(-x^2), x = z/sqrt 2
__device__ __forceinline__ bf_f2 gelu_erf_grad2(bf_f2 z2) {
    float x0, x1, d0, d1, q0, q1, r0, r1;
    bf_unpack2(bf_mul2(z2, bf_splat2(0.70710678118654752f)), x0, x1);
    const bf_f2 ax2 = bf_pack2(fabsf(x0), fabsf(x1));
    bf_unpack2(bf_fma2(bf_splat2(0.3275911f), ax2, bf_splat2(1.0f)), d0, d1);
    const bf_f2 t2 = bf_pack2(bf_rcp_approx(d0), bf_rcp_approx(d1));
    bf_f2 p2 = bf_fma2(bf_splat2(1.061405429f), t2, bf_splat2(-1.453152027f));
    p2 = bf_fma2(p2, t2, bf_splat2(1.421413741f));
    p2 = bf_fma2(p2, t2, bf_splat2(-0.284496736f));
    p2 = bf_fma2(p2, t2, bf_splat2(0.254829592f));
    bf_unpack2(bf_mul2(bf_mul2(ax2, ax2), bf_splat2(-1.4426950408889634f)), q0, q1);
    const float e0 = bf_ex2_approx(q0), e1 = bf_ex2_approx(q1);
    bf_unpack2(bf_fma2(bf_mul2(p2, t2), bf_pack2(-e0, -e1), bf_splat2(1.0f)), r0, r1);  // |erf|
    const bf_f2 cdf2 = bf_fma2(bf_pack2(copysignf(r0, x0), copysignf(r1, x1)), bf_splat2(0.5f), bf_splat2(0.5f));
    return bf_fma2(z2, bf_mul2(bf_pack2(e0, e1), bf_splat2(0.3989422804014327f)), cdf2);
}
__device__ __forceinline__ float gelu_erf_grad(float z) {
    const float cdf = 0.5f * (1.0f + erf_as(z * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * bf_ex2_approx(-0.72134752044448170f * z * z);  // exp(-z^2/2)/sqrt(2 pi)
    return fmaf(z, pdf, cdf);
}

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
    uint4 u;
    const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
    const __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
    u.x = *reinterpret_cast<const uint32_t*>(&p0), u.y = *reinterpret_cast<const uint32_t*>(&p1);
    u.z = *reinterpret_cast<const uint32_t*>(&p2), u.w = *reinterpret_cast<const uint32_t*>(&p3);
    return u;
}

template <bool POLY>  // POLY: polynomial GELU (BF_OPT_GELU_POLY = 1, default), else the erf form
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    bayes_gemm2_gelu_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                            const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_y,
                            const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t out_base = smem_base + kStages * STAGE_BYTES;
    uint8_t* const out_gen = smem_gen + kStages * STAGE_BYTES;
    const uint32_t bar_base = out_base + OUT_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    volatile uint32_t* const tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(out_gen + OUT_BYTES + 8 * (2 * kStages + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_z);
        tma_prefetch_desc(&map_y);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 2 * EPI_WARPS);  // leader's copy: epilogue warps of both CTAs
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    const int64_t n_items = p.S * p.i_pairs * p.j_tiles;
    const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t L = cluster_id; L < n_items; L += n_clusters) {
                const Item it = decode_item(p, L);
                const int i0 = it.i_pair * (2 * BLOCK_M) + (int)rank * BLOCK_M;
                const int j0 = it.j_blk * BLOCK_N + (int)rank * LOAD_N;
                for (int ks = 0; ks < p.k_steps; ++ks) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
                    tma_load_3d_2sm(a_dst, &map_a, full_bar(stage), ks * BLOCK_K, i0, it.s);
                    tma_load_3d_2sm(a_dst + A_BYTES, &map_b, full_bar(stage), ks * BLOCK_K, j0, it.s);
                    if (++stage == kStages) stage = 0, phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA, one thread) =====================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc(false, false, 2 * BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int64_t L = cluster_id; L < n_items; L += n_clusters, ++iter) {
                const int acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
                BF_GSTAMP(0, iter, 0);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                BF_GSTAMP(0, iter, 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int ks = 0; ks < p.k_steps; ++ks) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_src = smem_base + stage * STAGE_BYTES;
                    const uint32_t b_src = a_src + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_bf16_2sm(d_tmem, operand_desc<false>(a_src, k), operand_desc<false>(b_src, k), idesc,
                                      (ks > 0 || k > 0) ? 1u : 0u);
                    umma_commit_2sm(empty_bar(stage));
                    if (++stage == kStages) stage = 0, phase ^= 1u;
                }
                umma_commit_2sm(tfull_bar(acc));
                BF_GSTAMP(0, iter, 2);
            }
        }
    } else {
        // ===================== epilogue: 2 groups x 4 warps, own TMEM half =====================
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const bool store_thread = ((warp - 2) & 3) == 0 && lane == 0;
        // Staging: per group two sets of (z, y) HALF boxes [128 rows][32 columns] (64 B rows, 64B swizzle).  A half box
        // is computed into one set while the TMA stores of the other set drain: the store thread only ever waits for a
        // store issued a whole half box earlier (with one 64-column set per group the wait for the just-issued store
        // was 20 % of the tile time), and one barrier per half box is enough.
        const uint32_t grp_out = out_base + grp * 4 * HB_BYTES;
        uint8_t* const grp_out_gen = out_gen + grp * 4 * HB_BYTES;
        int iter = 0, n_hb = 0;
        for (int64_t L = cluster_id; L < n_items; L += n_clusters, ++iter) {
            const Item it = decode_item(p, L);
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int i0 = it.i_pair * (2 * BLOCK_M) + (int)rank * BLOCK_M, j0 = it.j_blk * BLOCK_N;
            if (store_thread) BF_GSTAMP(1 + grp, iter, 0);
            mbar_wait(tfull_bar(acc), acc_phase);
            if (store_thread) BF_GSTAMP(1 + grp, iter, 1);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            const int n_cols = (int64_t)j0 + BLOCK_N > p.J ? (int)(p.J - j0) : BLOCK_N;
            // this group's half boxes: columns c0 = 64 b + 32 hh, b = grp, grp + 2; the last one releases the accumulator
            int last_c0 = -1;
            for (int k = 0; k < 4; ++k) {
                const int c0 = (grp + 2 * (k >> 1)) * BOX_COLS + (k & 1) * HB_COLS;
                if (c0 < n_cols) last_c0 = c0;
            }
            if (last_c0 < 0) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
            }
            const float* const bias_row = p.bias + (int64_t)it.s * p.J + j0;
#pragma unroll 1
            for (int k = 0; k < 4; ++k) {
                const int c0 = (grp + 2 * (k >> 1)) * BOX_COLS + (k & 1) * HB_COLS;
                if (c0 >= n_cols) continue;
                uint32_t r[32];
                tmem_ld_32x32(t_acc + (uint32_t)c0, r);
                tmem_ld_wait();
                if (store_thread) BF_GSTAMP(1 + grp, iter, 2 + 3 * k);
                if (c0 == last_c0) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
                }
                const int set = n_hb++ & 1;
                uint8_t* const z_row = grp_out_gen + set * 2 * HB_BYTES + row * 64;
                uint8_t* const y_row = z_row + HB_BYTES;
#pragma unroll
                for (int t = 0; t < 4; ++t) {  // 8 columns = one 16-byte chunk of the bf16 row
                    float bv[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
                    if (c0 + 8 * t < n_cols) {  // N % 8 == 0: a chunk is wholly inside or wholly outside; same address in all lanes
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias_row + c0 + 8 * t));
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias_row + c0 + 8 * t) + 1);
                        bv[0] = b0.x, bv[1] = b0.y, bv[2] = b0.z, bv[3] = b0.w, bv[4] = b1.x, bv[5] = b1.y, bv[6] = b1.z, bv[7] = b1.w;
                    }
                    uint32_t zw[4], yw[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float z0, z1, y0, y1;
                        bf_unpack2(bf_add2(bf_pack2(__uint_as_float(r[8 * t + 2 * e]), __uint_as_float(r[8 * t + 2 * e + 1])),
                                           bf_pack2(bv[2 * e], bv[2 * e + 1])), z0, z1);
                        const __nv_bfloat162 zb = __floats2bfloat162_rn(z0, z1);
                        zw[e] = *reinterpret_cast<const uint32_t*>(&zb);
                        // gelu of the bf16-ROUNDED pre-activation: backward recomputes gelu' from the stored z
                        {
                            const bf_f2 zr2 = bf_pack2(__uint_as_float(zw[e] << 16), __uint_as_float(zw[e] & 0xffff0000u));
                            bf_unpack2(POLY ? gelu_poly2(zr2) : gelu_erf2(zr2), y0, y1);
                        }
                        const __nv_bfloat162 yb = __floats2bfloat162_rn(y0, y1);
                        yw[e] = *reinterpret_cast<const uint32_t*>(&yb);
                    }
                    const int sw = (t ^ ((row >> 1) & 3)) << 4;  // 64B swizzle: 16 B chunk index ^ bits 7..8 of the offset
                    *reinterpret_cast<uint4*>(z_row + sw) = make_uint4(zw[0], zw[1], zw[2], zw[3]);
                    *reinterpret_cast<uint4*>(y_row + sw) = make_uint4(yw[0], yw[1], yw[2], yw[3]);
                }
                fence_proxy_async();
                if (store_thread) BF_GSTAMP(1 + grp, iter, 3 + 3 * k);
                if (store_thread) tma_store_wait_read<0>();  // the other set's stores (a half box ago) have read smem
                if (store_thread) BF_GSTAMP(1 + grp, iter, 4 + 3 * k);
                named_bar_sync_dyn(1 + grp, 128);
                if (store_thread) {
                    const uint32_t st = grp_out + set * 2 * HB_BYTES;
                    tma_store_3d(&map_z, st, j0 + c0, i0, it.s);
                    tma_store_3d(&map_y, st + HB_BYTES, j0 + c0, i0, it.s);
                    tma_store_commit();
                }
            }
        }
        if (store_thread) tma_store_wait_all();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------ dgrad with the GELU derivative in its epilogue
// gz[s] = (gy[s] . w[s]) o gelu'(z[s]): the input gradient of the Linear that CONSUMES a = gelu(z), already multiplied
// by the activation's derivative, i.e. the gradient of the pre-activation z of the layer that PRODUCED a.  Replaces
// bf_linear_dgrad (linear.py:104's autograd) followed by the separate GELU' pass: the [S*M, N] gradient makes one trip
// through HBM instead of three (write da, read da + z, write gz -> read z, write gz).
//
// Kernel = the CTA-pair dgrad of bf_gemm_tc2.cu (A = gy K-major, B = w MN-major, UMMA 256 x 256 x 16) with the 8-warp /
// 2-group epilogue of the fused GELU forward.  The z tile travels like an operand: the TMA warp loads its four
// 128 x 64 boxes into dedicated buffers right after the tile's last operand stage (by then the previous tile's epilogue
// has released them), each epilogue group multiplies its boxes by gelu'(z) read from shared memory.
namespace dgelu {
using namespace act;

constexpr int kStages = 4;
constexpr int kThreads = 32 * (3 + EPI_WARPS);  // warp 0 operand TMA, warp 1 MMA, warp 2 z TMA, warps 3..10 epilogue
constexpr int Z_BYTES = BOXES * BOX_BYTES;             // 4 boxes of the tile's z rows: 64 KiB
constexpr int OUT_BYTES = EPI_GROUPS * BOX_BYTES;      // one staging box per epilogue group
constexpr int SMEM_BYTES = 1024 + kStages * STAGE_BYTES + Z_BYTES + OUT_BYTES + 256;

template <bool POLY, bool COLSUM>  // COLSUM: also emit the column sums of the result (the producer's bias gradient)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    bayes_gemm2_dgelu_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                             const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_out,
                             const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t z_base = smem_base + kStages * STAGE_BYTES;
    uint8_t* const z_gen = smem_gen + kStages * STAGE_BYTES;
    const uint32_t out_base = z_base + Z_BYTES;
    uint8_t* const out_gen = z_gen + Z_BYTES;
    const uint32_t bar_base = out_base + OUT_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
    auto zfull_bar = [&](int b) { return bar_base + 8u * (2 * kStages + 4 + b); };
    auto zempty_bar = [&](int b) { return bar_base + 8u * (2 * kStages + 4 + BOXES + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4 + 2 * BOXES);
    volatile uint32_t* const tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(out_gen + OUT_BYTES + 8 * (2 * kStages + 4 + 2 * BOXES));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_z);
        tma_prefetch_desc(&map_out);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 2 * EPI_WARPS);  // leader's copy: epilogue warps of both CTAs
        }
        for (int b = 0; b < BOXES; ++b) {
            mbar_init(zfull_bar(b), 1);   // own CTA: the TMA thread's arrive.expect_tx
            mbar_init(zempty_bar(b), 4);  // own CTA: the four warps of the group that reads box b
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    const int64_t n_items = p.S * p.i_pairs * p.j_tiles;
    const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs): operands, then this CTA's z boxes =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t L = cluster_id; L < n_items; L += n_clusters) {
                const Item it = decode_item(p, L);
                const int i0 = it.i_pair * (2 * BLOCK_M) + (int)rank * BLOCK_M;  // this CTA's rows of A / D
                const int j0 = it.j_blk * BLOCK_N;
                for (int ks = 0; ks < p.k_steps; ++ks) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
                    const uint32_t b_dst = a_dst + A_BYTES;
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
                    const int r0 = ks * BLOCK_K;
                    tma_load_3d_2sm(a_dst, &map_a, full_bar(stage), r0, i0, it.s);  // gy: K-major
#pragma unroll
                    for (int a = 0; a < LOAD_N / ATOM_MN; ++a)                      // w: MN-major, this CTA's half
                        tma_load_3d_2sm(b_dst + a * ATOM_BYTES, &map_b, full_bar(stage),
                                        j0 + (int)rank * LOAD_N + a * ATOM_MN, r0, it.s);
                    if (++stage == kStages) stage = 0, phase ^= 1u;
                }
            }
        }
    } else if (warp == 2) {
        // ===================== z loader (both CTAs, own rows): runs ahead of the epilogue by one tile ===============
        if (lane == 0) {
            uint32_t zphase = 0;
            for (int64_t L = cluster_id; L < n_items; L += n_clusters) {
                const Item it = decode_item(p, L);
                const int i0 = it.i_pair * (2 * BLOCK_M) + (int)rank * BLOCK_M;
                const int j0 = it.j_blk * BLOCK_N;
                int n_boxes = BOXES;
                if ((int64_t)j0 + BLOCK_N > p.J) n_boxes = (int)((p.J - j0 + BOX_COLS - 1) / BOX_COLS);
                for (int b = 0; b < BOXES; ++b) {
                    mbar_wait(zempty_bar(b), zphase ^ 1u);  // the previous tile's readers of this buffer are done
                    if (b < n_boxes) {
                        mbar_expect_tx(zfull_bar(b), BOX_BYTES);
                        tma_load_3d(z_base + b * BOX_BYTES, &map_z, zfull_bar(b), j0 + b * BOX_COLS, i0, it.s);
                    } else {
                        mbar_arrive(zfull_bar(b));  // box outside the matrix: nothing to load, keep the phases in step
                    }
                }
                zphase ^= 1u;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA, one thread) =====================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc(false, true, 2 * BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int64_t L = cluster_id; L < n_items; L += n_clusters, ++iter) {
                const int acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int ks = 0; ks < p.k_steps; ++ks) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_src = smem_base + stage * STAGE_BYTES;
                    const uint32_t b_src = a_src + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_bf16_2sm(d_tmem, operand_desc<false>(a_src, k), operand_desc<true>(b_src, k), idesc,
                                      (ks > 0 || k > 0) ? 1u : 0u);
                    umma_commit_2sm(empty_bar(stage));
                    if (++stage == kStages) stage = 0, phase ^= 1u;
                }
                umma_commit_2sm(tfull_bar(acc));
            }
        }
    } else {
        // ===================== epilogue: 2 groups x 4 warps, own TMEM half; group g takes boxes b % 2 == g ==========
        const int q = warp & 3;  // TMEM lane quarter this warp may touch (warps 3..6 and 7..10 each cover all four)
        const int grp = (warp - 3) >> 2;
        const int row = q * 32 + lane;
        const bool store_thread = ((warp - 3) & 3) == 0 && lane == 0;
        const uint32_t my_out = out_base + grp * BOX_BYTES;
        uint8_t* const o_row = out_gen + grp * BOX_BYTES + row * 128;
        int iter = 0;
        for (int64_t L = cluster_id; L < n_items; L += n_clusters, ++iter) {
            const Item it = decode_item(p, L);
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const uint32_t zphase = iter & 1;
            const int i0 = it.i_pair * (2 * BLOCK_M) + (int)rank * BLOCK_M, j0 = it.j_blk * BLOCK_N;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            int n_boxes = BOXES;
            if ((int64_t)j0 + BLOCK_N > p.J) n_boxes = (int)((p.J - j0 + BOX_COLS - 1) / BOX_COLS);
            int last_b = -1;
            for (int b = grp; b < n_boxes; b += EPI_GROUPS) last_b = b;
            if (last_b < 0) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
            }
#pragma unroll 1
            for (int b = grp; b < BOXES; b += EPI_GROUPS) {
                if (b >= n_boxes) {  // nothing to compute, but the z buffer's phases must advance
                    mbar_wait(zfull_bar(b), zphase);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(zempty_bar(b));
                    continue;
                }
                uint32_t r[2][32];
                tmem_ld_32x32(t_acc + (uint32_t)(b * BOX_COLS), r[0]);
                tmem_ld_32x32(t_acc + (uint32_t)(b * BOX_COLS + 32), r[1]);
                if (store_thread) tma_store_wait_read<0>();  // this group's previous store has read its staging box
                tmem_ld_wait();
                if (b == last_b) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
                }
                mbar_wait(zfull_bar(b), zphase);  // this tile's z box has landed
                named_bar_sync_dyn(1 + grp, 128);  // ... and the staging box is free
                const uint8_t* const z_row = z_gen + b * BOX_BYTES + row * 128;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {  // 8 columns = one 16-byte chunk of the bf16 row
                        const int ch = h * 4 + t;
                        const uint4 zz = *reinterpret_cast<const uint4*>(z_row + ((ch ^ (row & 7)) << 4));
                        const uint32_t zw[4] = {zz.x, zz.y, zz.z, zz.w};
                        uint32_t ow[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const bf_f2 g2 = bf_pack2(__uint_as_float(r[h][8 * t + 2 * e]), __uint_as_float(r[h][8 * t + 2 * e + 1]));
                            const bf_f2 z2 = bf_pack2(__uint_as_float(zw[e] << 16), __uint_as_float(zw[e] & 0xffff0000u));
                            float o0, o1;
                            bf_unpack2(bf_mul2(g2, POLY ? gelu_grad_poly2(z2) : gelu_erf_grad2(z2)), o0, o1);
                            const __nv_bfloat162 ob = __floats2bfloat162_rn(o0, o1);
                            ow[e] = *reinterpret_cast<const uint32_t*>(&ob);
                        }
                        *reinterpret_cast<uint4*>(o_row + ((ch ^ (row & 7)) << 4)) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(zempty_bar(b));  // this warp's reads of the z box are done
                fence_proxy_async();
                named_bar_sync_dyn(1 + grp, 128);
                if (store_thread) {
                    tma_store_3d(&map_out, my_out, j0 + b * BOX_COLS, i0, it.s);
                    tma_store_commit();
                }
                if (COLSUM) {
                    // column sums of the staged (bf16-rounded) box: warp w of the group takes columns 16 w .. +15 (8 words
                    // of a row), lane l word l % 8 of the rows r = l / 8 (mod 4); rows beyond M were computed from
                    // zero-filled operands and z, so they hold zeros.  The next box's barrier comes after these reads.
                    const int gw = (warp - 3) & 3, wd = lane & 7;
                    const uint8_t* const base = out_gen + grp * BOX_BYTES + (wd & 3) * 4;  // word wd % 4 of 16 B chunk 2 gw + wd / 4
                    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr) {
                        const int r_ = 4 * rr + (lane >> 3);
                        const uint32_t v = *reinterpret_cast<const uint32_t*>(base + r_ * 128 + (((gw * 2 + (wd >> 2)) ^ (r_ & 7)) << 4));
                        s0 += __uint_as_float(v << 16);
                        s1 += __uint_as_float(v & 0xffff0000u);
                    }
                    s0 += __shfl_xor_sync(0xffffffffu, s0, 8), s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
                    s0 += __shfl_xor_sync(0xffffffffu, s0, 16), s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                    const int64_t col = (int64_t)j0 + b * BOX_COLS + gw * 16 + wd * 2;
                    if (lane < 8 && col < p.J) {  // J % 8 == 0: the pair is inside or outside together
                        // fire-and-forget adds (a load + add + store would put an L2 round trip on the epilogue's critical
                        // path); this thread is the only one that ever touches these two addresses and its adds to one
                        // address are applied in program order, so the sum is still deterministic
                        float* const dst = p.col_partial + ((int64_t)blockIdx.x * p.S + it.s) * p.J + col;
                        atomicAdd(dst, s0);
                        atomicAdd(dst + 1, s1);
                    }
                }
            }
        }
        if (store_thread) tma_store_wait_all();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    }
}

}  // namespace dgelu

// ------------------------------------------------------------------ backward: gz = gy * gelu'(z), db = colsum(gz)
constexpr int kBwThreads = 256, kBwWarps = 8, kBwCols = 256;  // 32 lanes x 8 bf16 columns

__global__ void __launch_bounds__(kBwThreads) gelu_bwd_bias_grad_kernel(const __nv_bfloat16* __restrict__ gy,
                                                                        const __nv_bfloat16* __restrict__ z,
                                                                        __nv_bfloat16* __restrict__ gz,
                                                                        float* __restrict__ db,
                                                                        float* __restrict__ partial,
                                                                        unsigned int* __restrict__ counters, int64_t M,
                                                                        int64_t N, int64_t rows_per_slab) {
    const int s = blockIdx.y, slab = blockIdx.z, n_slabs = gridDim.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * kBwCols + lane * 8;
    const int64_t base = (int64_t)s * M * N;
    const int64_t m_lo = (int64_t)slab * rows_per_slab;
    const int64_t m_hi = m_lo + rows_per_slab < M ? m_lo + rows_per_slab : M;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    if (c0 < N) {  // N % 8 == 0
        // one 16-byte chunk: 8 (gy, z) pairs -> 8 products (packed fp32 math), accumulated per column
        auto chunk = [&](const uint4& g, const uint4& zz, float (&o)[8]) {
            const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, zw[4] = {zz.x, zz.y, zz.z, zz.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bf_f2 g2 = bf_pack2(__uint_as_float(gw[i] << 16), __uint_as_float(gw[i] & 0xffff0000u));
                const bf_f2 z2 = bf_pack2(__uint_as_float(zw[i] << 16), __uint_as_float(zw[i] & 0xffff0000u));
                bf_unpack2(bf_mul2(g2, gelu_erf_grad2(z2)), o[2 * i], o[2 * i + 1]);
                acc[2 * i] += o[2 * i], acc[2 * i + 1] += o[2 * i + 1];
            }
        };
        int64_t m = m_lo + warp;
        for (; m + kBwWarps < m_hi; m += 2 * kBwWarps) {  // 2 rows in flight per warp
            const int64_t o0 = base + m * N + c0, o1 = o0 + (int64_t)kBwWarps * N;
            const uint4 g0 = __ldcs(reinterpret_cast<const uint4*>(gy + o0)), z0 = __ldcs(reinterpret_cast<const uint4*>(z + o0));
            const uint4 g1 = __ldcs(reinterpret_cast<const uint4*>(gy + o1)), z1 = __ldcs(reinterpret_cast<const uint4*>(z + o1));
            float o[8];
            chunk(g0, z0, o);
            __stcs(reinterpret_cast<uint4*>(gz + o0), pack8_bf16(o));
            chunk(g1, z1, o);
            __stcs(reinterpret_cast<uint4*>(gz + o1), pack8_bf16(o));
        }
        for (; m < m_hi; m += kBwWarps) {
            const int64_t o0 = base + m * N + c0;
            const uint4 g0 = __ldcs(reinterpret_cast<const uint4*>(gy + o0)), z0 = __ldcs(reinterpret_cast<const uint4*>(z + o0));
            float o[8];
            chunk(g0, z0, o);
            __stcs(reinterpret_cast<uint4*>(gz + o0), pack8_bf16(o));
        }
    }
    // NOTE: db sums the fp32 products before they are rounded to bf16 for gz (more accurate than summing gz)
    __shared__ float red[kBwWarps][kBwCols];
    __shared__ bool is_last;
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    const int64_t cb = blockIdx.x, n_cb = gridDim.x;
    {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kBwWarps; ++w) t += red[w][threadIdx.x];
        partial[(((int64_t)s * n_cb + cb) * n_slabs + slab) * kBwCols + threadIdx.x] = t;
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
    {
        const volatile float* pp = partial + ((int64_t)s * n_cb + cb) * n_slabs * kBwCols + threadIdx.x;
        float t = 0.0f;
        for (int k = 0; k < n_slabs; ++k) t += pp[(int64_t)k * kBwCols];
        const int64_t c = cb * kBwCols + threadIdx.x;
        if (c < N) db[(int64_t)s * N + c] = t;
    }
    if (threadIdx.x == 0) counters[s * n_cb + cb] = 0u;
}

inline void bw_grid(int64_t S, int64_t M, int64_t N, int& n_cb, int& n_slabs, int64_t& rps) {
    n_cb = (int)((N + kBwCols - 1) / kBwCols);
    const int64_t target = (int64_t)bf_num_sms() * 4;
    int64_t slabs = target / (S * n_cb);
    const int64_t max_slabs = (M + 63) / 64;
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    if (slabs > 65535) slabs = 65535;
    rps = (M + slabs - 1) / slabs;
    n_slabs = (int)((M + rps - 1) / rps);
}

}  // namespace act

extern "C" int bf_linear_fwd_gelu_supported(int64_t S, int64_t M, int64_t N, int64_t K) {
    if (N % 8 != 0 || K % 8 != 0 || M < 256) return 0;
    const int64_t pair_tiles = S * ((M + 255) / 256) * ((N + 255) / 256);
    return pair_tiles >= bf_num_sms() / 2 ? 1 : 0;
}

// z[s] = x[s] w[s]^T + bias[s] (bf16), y[s] = gelu(z[s]) (bf16)
extern "C" int bf_linear_fwd_gelu(const void* x, const void* w, const float* bias, void* z, void* y, int64_t S,
                                  int64_t M, int64_t N, int64_t K, void* stream) {
    using namespace act;
    BF_CHECK_ARG(x && w && bias && z && y, "null pointer (the fused GELU forward needs a bias)");
    BF_CHECK_ARG(S >= 1 && M >= 1 && N >= 1 && K >= 1, "S, M, N, K must be >= 1");
    BF_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "bf16 path needs K % 8 == 0 and N % 8 == 0");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap ma, mb, mz, my;
    int rc;
    if ((rc = tc::encode_map(&ma, x, S, M, K, BLOCK_M))) return rc;
    if ((rc = tc::encode_map(&mb, w, S, N, K, LOAD_N))) return rc;
    if ((rc = tc::encode_map(&mz, z, S, M, N, BLOCK_M, false, HB_COLS))) return rc;  // 64 B half boxes, 64B swizzle
    if ((rc = tc::encode_map(&my, y, S, M, N, BLOCK_M, false, HB_COLS))) return rc;
    Params p{};
    p.S = S, p.I = M, p.J = N, p.R = K;
    p.i_pairs = tc::cdiv(M, 2 * BLOCK_M), p.j_tiles = tc::cdiv(N, BLOCK_N), p.k_steps = tc::cdiv(K, tc::BLOCK_K);
    p.bias = bias;
    auto* const kernel = bf_option(BF_OPT_GELU_POLY) ? bayes_gemm2_gelu_kernel<true> : bayes_gemm2_gelu_kernel<false>;
    BF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int64_t n_items = p.S * p.i_pairs * p.j_tiles;
    const int64_t pairs = bf_num_sms() / 2;
    const int grid = 2 * (int)(n_items < pairs ? n_items : pairs);
    kernel<<<grid, kThreads, SMEM_BYTES, st>>>(ma, mb, mz, my, p);
    BF_LAUNCH_OK();
    return 0;
}

extern "C" int64_t bf_gelu_bwd_bias_grad_workspace_bytes(int64_t S, int64_t M, int64_t N) {
    S = S < 1 ? 1 : S, M = M < 1 ? 1 : M, N = N < 1 ? 1 : N;
    int n_cb, n_slabs;
    int64_t rps;
    act::bw_grid(S, M, N, n_cb, n_slabs, rps);
    return ((S * n_cb * 4 + 255) / 256) * 256 + S * n_cb * (int64_t)n_slabs * act::kBwCols * 4;
}

// gz = gy * gelu'(z) (bf16), db[s][n] = sum_m gz[s][m][n] (fp32).  workspace zero-filled once.
extern "C" int bf_gelu_bwd_bias_grad(const void* gy, const void* z, void* gz, float* db, int64_t S, int64_t M,
                                     int64_t N, void* workspace, void* stream) {
    using namespace act;
    BF_CHECK_ARG(gy && z && gz && db && workspace, "null pointer");
    BF_CHECK_ARG(S >= 1 && M >= 1 && N >= 8 && N % 8 == 0, "needs N % 8 == 0");
    BF_CHECK_ARG(((reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(gz)) & 15u) == 0,
                 "gy, z, gz must be 16 B aligned");
    int n_cb, n_slabs;
    int64_t rps;
    bw_grid(S, M, N, n_cb, n_slabs, rps);
    const int64_t cnt = ((S * n_cb * 4 + 255) / 256) * 256;
    unsigned int* counters = reinterpret_cast<unsigned int*>(workspace);
    float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + cnt);
    dim3 grid((unsigned)n_cb, (unsigned)S, (unsigned)n_slabs);
    gelu_bwd_bias_grad_kernel<<<grid, kBwThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(gy), reinterpret_cast<const __nv_bfloat16*>(z),
        reinterpret_cast<__nv_bfloat16*>(gz), db, partial, counters, M, N, rps);
    BF_LAUNCH_OK();
    return 0;
}

extern "C" int bf_linear_dgrad_gelu_supported(int64_t S, int64_t M, int64_t N, int64_t K) {
    // same occupancy rule as the fused forward: enough 256 x 256 tiles of the [M, K] result for every SM pair
    return bf_linear_fwd_gelu_supported(S, M, K, N);
}

namespace act {
// dbias[i] = sum over the per-block rows of the partial column sums, in block order (deterministic)
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                            int64_t n, int blocks) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float acc = 0.0f;
    for (int c = 0; c < blocks; ++c) acc += partial[(int64_t)c * n + i];
    out[i] = acc;
}
static int dgelu_grid(int64_t S, int64_t M, int64_t K) {
    const int64_t n_items = S * tc::cdiv(M, 2 * BLOCK_M) * tc::cdiv(K, BLOCK_N);
    const int64_t pairs = bf_num_sms() / 2;
    return 2 * (int)(n_items < pairs ? n_items : pairs);
}
}  // namespace act

extern "C" int64_t bf_linear_dgrad_gelu_bias_workspace_bytes(int64_t S, int64_t M, int64_t K) {
    S = S < 1 ? 1 : S, M = M < 1 ? 1 : M, K = K < 1 ? 1 : K;
    return (int64_t)act::dgelu_grid(S, M, K) * S * K * 4;
}

// gz[s] = (gy[s] . w[s]) o gelu'(z[s])   gy [S,M,N], w [S,N,K], z / gz [S,M,K], all bf16;
// dbias (optional, with its workspace): dbias[s][k] = sum_m gz[s][m][k], fp32 -- the bias gradient of the layer that made z
extern "C" int bf_linear_dgrad_gelu_bias(const void* gy, const void* w, const void* z, void* gz, float* dbias, void* workspace,
                                         int64_t S, int64_t M, int64_t N, int64_t K, void* stream) {
    namespace dg = act::dgelu;
    BF_CHECK_ARG(gy && w && z && gz, "null pointer");
    BF_CHECK_ARG(!dbias || workspace, "the bias gradient needs its workspace (bf_linear_dgrad_gelu_bias_workspace_bytes)");
    BF_CHECK_ARG(S >= 1 && M >= 1 && N >= 1 && K >= 1, "S, M, N, K must be >= 1");
    BF_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "bf16 path needs K % 8 == 0 and N % 8 == 0");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap ma, mb, mz, mo;
    int rc;
    if ((rc = tc::encode_map(&ma, gy, S, M, N, act::BLOCK_M))) return rc;  // A: K-major over r = N
    if ((rc = tc::encode_map(&mb, w, S, N, K, tc::BLOCK_K))) return rc;    // B: MN-major, rows = r = N, cols = K
    if ((rc = tc::encode_map(&mz, z, S, M, K, act::BLOCK_M))) return rc;
    if ((rc = tc::encode_map(&mo, gz, S, M, K, act::BLOCK_M))) return rc;
    act::Params p{};
    p.S = S, p.I = M, p.J = K, p.R = N;
    p.i_pairs = tc::cdiv(M, 2 * act::BLOCK_M), p.j_tiles = tc::cdiv(K, act::BLOCK_N), p.k_steps = tc::cdiv(N, tc::BLOCK_K);
    const int grid = act::dgelu_grid(S, M, K);
    if (dbias) {
        p.col_partial = static_cast<float*>(workspace);
        BF_CUDA_OK(cudaMemsetAsync(workspace, 0, (size_t)grid * S * K * 4, st));
    }
    const bool poly = bf_option(BF_OPT_GELU_POLY) != 0;
    auto* const kernel = dbias ? (poly ? dg::bayes_gemm2_dgelu_kernel<true, true> : dg::bayes_gemm2_dgelu_kernel<false, true>)
                               : (poly ? dg::bayes_gemm2_dgelu_kernel<true, false> : dg::bayes_gemm2_dgelu_kernel<false, false>);
    BF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dg::SMEM_BYTES));
    kernel<<<grid, dg::kThreads, dg::SMEM_BYTES, st>>>(ma, mb, mz, mo, p);
    BF_LAUNCH_OK();
    if (dbias) {
        const int64_t n = S * K;
        act::colsum_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p.col_partial, dbias, n, grid);
        BF_LAUNCH_OK();
    }
    return 0;
}

extern "C" int bf_linear_dgrad_gelu(const void* gy, const void* w, const void* z, void* gz, int64_t S, int64_t M,
                                    int64_t N, int64_t K, void* stream) {
    return bf_linear_dgrad_gelu_bias(gy, w, z, gz, nullptr, nullptr, S, M, N, K, stream);
}
