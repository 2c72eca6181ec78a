// CTA-pair (cta_group::2) version of the S-sample Linear contractions of bf_gemm_tc.cu.
//
// Same contract (fwd: y = x w^T + b, dgrad: dx = gy w; replaces F.linear,
// bayeformers/nn/layers/linear.py:104, and the mm of its autograd), different
// machine mapping: a cluster of two CTAs on the two SMs of one TPC computes a
// 256 x 256 output tile with ONE tcgen05.mma.cta_group::2 stream (UMMA M = 256).
// Each CTA loads only ITS half of A (128 rows) and ITS half of B (128 of the 256
// B rows) -- 32 KiB per 64-deep k-step instead of the 48 KiB a lone 128 x 256 CTA
// needs -- so the operand feed per SM drops by a third and the ring gets 6 stages.
// ncu showed the single-CTA kernel limited by exactly that feed (profiles/README.md).
//
//   TMA   : both CTAs, own halves into own smem, complete_tx on the LEADER's `full` barrier
//   MMA   : leader only; commits multicast to both CTAs' `empty` / `tmem_full` barriers
//   TMEM  : rows [128 r, 128 r + 128) of the tile live in CTA r (2 x 256 columns, double-buffered)
//   epilogue: every CTA drains its own TMEM half (tcgen05.ld -> bias -> swizzled smem -> TMA store)
//             and arrives on the leader's `tmem_empty` barrier
#include <type_traits>

#include "bf_tc.cuh"

namespace tc2 {
using namespace tc;

constexpr int BLOCK_M = 128;   // rows per CTA (UMMA M = 256 over the pair)
constexpr int BLOCK_N = 256;   // tile columns (UMMA N)
constexpr int LOAD_N = 128;    // B rows each CTA loads
constexpr int kStages = 6;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int B_BYTES = LOAD_N * BLOCK_K * 2;   // 16 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int OUT_BOX_BYTES = BLOCK_M * 128;
constexpr int EPI_WARPS = 4;
constexpr int kThreads = 32 * (2 + EPI_WARPS);
constexpr int TMEM_COLS = 2 * BLOCK_N;
constexpr int BIAS_BYTES = BLOCK_N * 4;  // this tile's bias slice, staged once per tile
constexpr int SMEM_BYTES = 1024 + kStages * STAGE_BYTES + 2 * OUT_BOX_BYTES + 256 + BIAS_BYTES;

struct Params {
    int64_t S, I, J, R;
    int i_pairs, j_tiles, k_steps;  // i_pairs: 256-row tile rows
    const float* bias;
    int accumulate;  // 1: the result tile is ADDED to D (TMA reduce-add) instead of stored
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

template <bool A_MN, bool B_MN, bool OUT_F32, bool HAS_BIAS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    bayes_gemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const __grid_constant__ CUtensorMap map_out, const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t out_base = smem_base + kStages * STAGE_BYTES;
    uint8_t* const out_gen = smem_gen + kStages * STAGE_BYTES;
    const uint32_t bar_base = out_base + 2 * OUT_BOX_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    volatile uint32_t* const tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(out_gen + 2 * OUT_BOX_BYTES + 8 * (2 * kStages + 4));

    float* const bias_gen = reinterpret_cast<float*>(out_gen + 2 * OUT_BOX_BYTES + 256);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();  // 0 = leader
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);   // used in the leader only: its own arrive.expect_tx, bytes from both CTAs
            mbar_init(empty_bar(s), 1);  // multicast commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);               // multicast commit
            mbar_init(tempty_bar(a), 2 * EPI_WARPS);  // used in the leader only: epilogue warps of both CTAs
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / complete_tx
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
                const int i0 = it.i_pair * (2 * BLOCK_M) + (int)rank * BLOCK_M;  // this CTA's rows of A / D
                const int j0 = it.j_blk * BLOCK_N + (int)rank * LOAD_N;          // this CTA's rows of B
                for (int ks = 0; ks < p.k_steps; ++ks) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
                    const uint32_t b_dst = a_dst + A_BYTES;
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
                    const int r0 = ks * BLOCK_K;
                    if (A_MN) {
#pragma unroll
                        for (int a = 0; a < BLOCK_M / ATOM_MN; ++a)
                            tma_load_3d_2sm(a_dst + a * ATOM_BYTES, &map_a, full_bar(stage), i0 + a * ATOM_MN, r0, it.s);
                    } else {
                        tma_load_3d_2sm(a_dst, &map_a, full_bar(stage), r0, i0, it.s);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int a = 0; a < LOAD_N / ATOM_MN; ++a)
                            tma_load_3d_2sm(b_dst + a * ATOM_BYTES, &map_b, full_bar(stage), j0 + a * ATOM_MN, r0, it.s);
                    } else {
                        tma_load_3d_2sm(b_dst, &map_b, full_bar(stage), r0, j0, it.s);
                    }
                    if (++stage == kStages) stage = 0, phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA, one thread) =====================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc(A_MN, B_MN, 2 * BLOCK_M, BLOCK_N);
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
                        umma_bf16_2sm(d_tmem, operand_desc<A_MN>(a_src, k), operand_desc<B_MN>(b_src, k), idesc,
                                      (ks > 0 || k > 0) ? 1u : 0u);
                    umma_commit_2sm(empty_bar(stage));  // frees this stage in both CTAs
                    if (++stage == kStages) stage = 0, phase ^= 1u;
                }
                umma_commit_2sm(tfull_bar(acc));  // accumulator complete -> both epilogues
            }
        }
    } else {
        // ===================== epilogue warps (both CTAs, own TMEM half) =====================
        using OutT = typename std::conditional<OUT_F32, float, __nv_bfloat16>::type;
        constexpr int BOX_COLS = 128 / (int)sizeof(OutT);
        constexpr int LDS_PER_BOX = BOX_COLS / 32;
        constexpr int BOXES = BLOCK_N / BOX_COLS;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const bool store_thread = (warp == 2 && lane == 0);
        int iter = 0;
        uint32_t box_count = 0;
        for (int64_t L = cluster_id; L < n_items; L += n_clusters, ++iter) {
            const Item it = decode_item(p, L);
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int i0 = it.i_pair * (2 * BLOCK_M) + (int)rank * BLOCK_M, j0 = it.j_blk * BLOCK_N;
            if (HAS_BIAS) {
                // all 128 epilogue threads need the same 256 bias values: fetch them once per tile, before the
                // accumulator wait (the previous tile's last named barrier already ordered the old reads)
                const int et = threadIdx.x - 64;  // 0..127
#pragma unroll
                for (int u = 0; u < BLOCK_N / (EPI_WARPS * 32); ++u) {
                    const int64_t jc = (int64_t)j0 + et + u * (EPI_WARPS * 32);
                    bias_gen[et + u * (EPI_WARPS * 32)] = jc < p.J ? __ldg(p.bias + (int64_t)it.s * p.J + jc) : 0.0f;
                }
            }
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            int n_boxes = BOXES;
            if ((int64_t)j0 + BLOCK_N > p.J) n_boxes = (int)((p.J - j0 + BOX_COLS - 1) / BOX_COLS);
#pragma unroll 1
            for (int b = 0; b < n_boxes; ++b, ++box_count) {
                const uint32_t buf = box_count & 1u;
                uint32_t r[LDS_PER_BOX][32];
#pragma unroll
                for (int h = 0; h < LDS_PER_BOX; ++h) tmem_ld_32x32(t_acc + (uint32_t)(b * BOX_COLS + h * 32), r[h]);
                if (store_thread) tma_store_wait_read<1>();
                tmem_ld_wait();
                if (b == n_boxes - 1) {  // this warp has drained the accumulator: tell the leader's MMA thread
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
                }
                named_bar_sync<1, EPI_WARPS * 32>();
                uint8_t* const my_row = out_gen + buf * OUT_BOX_BYTES + row * 128;
#pragma unroll
                for (int h = 0; h < LDS_PER_BOX; ++h) {
                    float v[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[h][e]);
                    if (HAS_BIAS) {
                        const float4* bp = reinterpret_cast<const float4*>(bias_gen + b * BOX_COLS + h * 32);
#pragma unroll
                        for (int e4 = 0; e4 < 8; ++e4) {  // same address in every lane: shared-memory broadcast
                            const float4 bv = bp[e4];
                            v[e4 * 4 + 0] += bv.x, v[e4 * 4 + 1] += bv.y;
                            v[e4 * 4 + 2] += bv.z, v[e4 * 4 + 3] += bv.w;
                        }
                    }
                    if (OUT_F32) {
#pragma unroll
                        for (int ch = 0; ch < 8; ++ch)
                            *reinterpret_cast<float4*>(my_row + ((ch ^ (row & 7)) << 4)) =
                                make_float4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]);
                    } else {
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int ch = h * 4 + t;
                            const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * t + 0], v[8 * t + 1]);
                            const __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * t + 2], v[8 * t + 3]);
                            const __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * t + 4], v[8 * t + 5]);
                            const __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * t + 6], v[8 * t + 7]);
                            uint4 u;
                            u.x = *reinterpret_cast<const uint32_t*>(&p0);
                            u.y = *reinterpret_cast<const uint32_t*>(&p1);
                            u.z = *reinterpret_cast<const uint32_t*>(&p2);
                            u.w = *reinterpret_cast<const uint32_t*>(&p3);
                            *reinterpret_cast<uint4*>(my_row + ((ch ^ (row & 7)) << 4)) = u;
                        }
                    }
                }
                fence_proxy_async();
                named_bar_sync<1, EPI_WARPS * 32>();
                if (store_thread) {
                    if (p.accumulate) tma_reduce_add_3d(&map_out, out_base + buf * OUT_BOX_BYTES, j0 + b * BOX_COLS, i0, it.s);
                    else tma_store_3d(&map_out, out_base + buf * OUT_BOX_BYTES, j0 + b * BOX_COLS, i0, it.s);
                    tma_store_commit();
                }
            }
        }
        if (store_thread) tma_store_wait_all();
    }

    // the peer must keep its shared memory / TMEM alive until the leader's last MMA has retired
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    }
}

template <bool A_MN, bool B_MN, bool OUT_F32, bool HAS_BIAS>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const Params& p,
                  cudaStream_t st) {
    auto kern = bayes_gemm2_kernel<A_MN, B_MN, OUT_F32, HAS_BIAS>;
    BF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int64_t n_items = p.S * p.i_pairs * p.j_tiles;
    const int64_t pairs = bf_num_sms() / 2;
    const int grid = 2 * (int)(n_items < pairs ? n_items : pairs);
    kern<<<grid, kThreads, SMEM_BYTES, st>>>(ma, mb, mo, p);  // cluster shape comes from __cluster_dims__
    BF_LAUNCH_OK();
    return 0;
}

template <bool A_MN, bool B_MN>
static int launch_out(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const Params& p,
                      bool out_f32, cudaStream_t st) {
    const bool bias = p.bias != nullptr;
    if (out_f32)
        return bias ? launch<A_MN, B_MN, true, true>(ma, mb, mo, p, st)
                    : launch<A_MN, B_MN, true, false>(ma, mb, mo, p, st);
    return bias ? launch<A_MN, B_MN, false, true>(ma, mb, mo, p, st)
                : launch<A_MN, B_MN, false, false>(ma, mb, mo, p, st);
}

}  // namespace tc2

// y[s] = x[s] w[s]^T + bias[s]   (CTA-pair kernel)
int bf_linear_fwd_bf16_2cta(const void* x, const void* w, const float* bias, void* y, int64_t S, int64_t M, int64_t N,
                            int64_t K, int32_t y_dtype, cudaStream_t st) {
    using namespace tc2;
    CUtensorMap ma, mb, mo;
    int rc;
    if ((rc = tc::encode_map(&ma, x, S, M, K, BLOCK_M))) return rc;
    if ((rc = tc::encode_map(&mb, w, S, N, K, LOAD_N))) return rc;
    if ((rc = tc::encode_map(&mo, y, S, M, N, BLOCK_M, y_dtype == BF_F32))) return rc;
    Params p{};
    p.S = S, p.I = M, p.J = N, p.R = K;
    p.i_pairs = tc::cdiv(M, 2 * BLOCK_M), p.j_tiles = tc::cdiv(N, BLOCK_N), p.k_steps = tc::cdiv(K, tc::BLOCK_K);
    p.bias = bias;
    return launch_out<false, false>(ma, mb, mo, p, y_dtype == BF_F32, st);
}

// dx[s] = gy[s] w[s]   (CTA-pair kernel)
int bf_linear_dgrad_bf16_2cta(const void* gy, const void* w, void* dx, int64_t S, int64_t M, int64_t N, int64_t K,
                              int32_t dx_dtype, int accumulate, cudaStream_t st) {
    using namespace tc2;
    CUtensorMap ma, mb, mo;
    int rc;
    if ((rc = tc::encode_map(&ma, gy, S, M, N, BLOCK_M))) return rc;     // A: K-major over r = N
    if ((rc = tc::encode_map(&mb, w, S, N, K, tc::BLOCK_K))) return rc;  // B: MN-major, rows = r = N, cols = K
    if ((rc = tc::encode_map(&mo, dx, S, M, K, BLOCK_M, dx_dtype == BF_F32))) return rc;
    Params p{};
    p.S = S, p.I = M, p.J = K, p.R = N;
    p.i_pairs = tc::cdiv(M, 2 * BLOCK_M), p.j_tiles = tc::cdiv(K, BLOCK_N), p.k_steps = tc::cdiv(N, tc::BLOCK_K);
    p.accumulate = accumulate;
    return launch_out<false, true>(ma, mb, mo, p, dx_dtype == BF_F32, st);
}
