// S-sample Linear contractions on 5th-gen tensor cores (sm_100a):
// TMA loads (cp.async.bulk.tensor, 128B swizzle) -> 4-stage shared-memory ring ->
// tcgen05.mma (bf16 x bf16 -> fp32 in TMEM, double-buffered accumulator) ->
// tcgen05.ld epilogue -> swizzled shared-memory staging -> TMA store.
// Warp-specialised (1 TMA warp, 1 MMA warp, 4 epilogue warps), persistent, one
// CTA per SM, 128 x 256 output tiles.
//
// Replaces F.linear (bayeformers/nn/layers/linear.py:104) and the mm calls of its
// autograd for all S Monte-Carlo samples at once.  One kernel, three contractions:
//
//     D[s][i][j] = sum_r A_s(i, r) * B_s(j, r)
//
//   fwd   : D = y  [M,N]   A = x  (K-major)   B = w  (K-major)   r = K   (+ bias)
//   dgrad : D = dx [M,K]   A = gy (K-major)   B = w  (MN-major)  r = N
//   wgrad : D = dW [N,K]   A = gy (MN-major)  B = x  (MN-major)  r = M   (unfused form, fp32 out)
//
// "K-major" = the reduction index is the contiguous one in memory; "MN-major" =
// the output index is contiguous, so no transposed copy is ever materialised.
// The variational (fused) weight gradient lives in bf_wgrad_tc.cu.
#include <cstdlib>
#include <type_traits>

#include "bf_tc.cuh"

namespace tc {

constexpr int BLOCK_M = 128;  // UMMA M (TMEM lanes)
constexpr int BLOCK_N = 256;  // UMMA N (TMEM columns per accumulator)
constexpr int kStages = 4;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;  // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int OUT_BOX_BYTES = BLOCK_M * 128;    // one TMA-store box: 128 rows x 128 B
constexpr int EPI_WARPS = 4;
constexpr int kThreads = 32 * (2 + EPI_WARPS);  // warp0 TMA, warp1 MMA, warps 2..5 epilogue
constexpr int TMEM_COLS = 2 * BLOCK_N;          // double-buffered accumulator = all 512 columns
constexpr int BIAS_BYTES = BLOCK_N * 4;  // this tile's bias slice, staged once per tile
constexpr int SMEM_BYTES = 1024 /*align slack*/ + kStages * STAGE_BYTES + 2 * OUT_BOX_BYTES + 256 + BIAS_BYTES;

struct Params {
    int64_t S, I, J, R;  // batch, rows of D, cols of D, reduction length
    int i_tiles, j_tiles, k_steps;
    const float* bias;  // [S][J] or null
    int accumulate;     // 1: the result tile is ADDED to D (TMA reduce-add) instead of stored
    // 1: D = A B.  3: split-precision ("fp32x3") contraction of fp32 operands given as bf16 (hi, lo) pairs,
    //    D = A_hi B_hi + A_hi B_lo + A_lo B_hi, three passes over the reduction into ONE fp32 TMEM accumulator
    //    (the dropped A_lo B_lo term is 2^-16 of a product): reference-precision results on the tensor cores
    int passes;
    // reduction chunks (>= 1).  The tensor core's fp32 accumulator truncates: ~3e-8 relative per accumulated MMA, which
    // the split-precision mode cannot afford over thousands of k-steps.  With r_chunks > 1 a work item covers
    // `ks_per_chunk` k-steps only and its tile is ADDED to D in global memory (TMA reduce-add, a properly rounded fp32
    // add in L2); D must be zero-filled by the caller, the bias rides on chunk 0.
    int r_chunks, ks_per_chunk;
};

struct Item {
    int s, i_blk, j_blk, chunk;
};
__device__ __forceinline__ Item decode_item(const Params& p, int64_t L) {
    Item it;
    it.j_blk = (int)(L % p.j_tiles);
    int64_t q = L / p.j_tiles;
    it.i_blk = (int)(q % p.i_tiles);
    q /= p.i_tiles;
    it.chunk = (int)(q % p.r_chunks);
    it.s = (int)(q / p.r_chunks);
    return it;
}
// k-steps [first, first + count) of the reduction handled by chunk `c`
__device__ __forceinline__ void chunk_range(const Params& p, int c, int& first, int& count) {
    first = c * p.ks_per_chunk;
    count = p.k_steps - first < p.ks_per_chunk ? p.k_steps - first : p.ks_per_chunk;
}

template <bool A_MN, bool B_MN, bool OUT_F32, bool HAS_BIAS>
__global__ void __launch_bounds__(kThreads, 1)
    bayes_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_a_lo,
                      const __grid_constant__ CUtensorMap map_b_lo, const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B swizzle needs 1024 B alignment
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

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
        if (p.passes > 1) {
            tma_prefetch_desc(&map_a_lo);
            tma_prefetch_desc(&map_b_lo);
        }
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

    const int64_t n_items = p.S * p.r_chunks * p.i_tiles * p.j_tiles;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t L = blockIdx.x; L < n_items; L += gridDim.x) {
                const Item it = decode_item(p, L);
                const int i0 = it.i_blk * BLOCK_M, j0 = it.j_blk * BLOCK_N;
                int ks0, ks_n;
                chunk_range(p, it.chunk, ks0, ks_n);
                for (int pass = 0; pass < p.passes; ++pass) {
                    // split precision: the two small cross terms first, (A_hi, B_hi) last -- the accumulator's
                    // truncation error is relative to its magnitude, so only the last third of the MMAs pays it in full
                    //   pass 0: (A_lo, B_hi)   pass 1: (A_hi, B_lo)   pass 2: (A_hi, B_hi);   passes == 1: (A, B)
                    const CUtensorMap* const ma = (p.passes > 1 && pass == 0) ? &map_a_lo : &map_a;
                    const CUtensorMap* const mb = (p.passes > 1 && pass == 1) ? &map_b_lo : &map_b;
                    for (int ks = ks0; ks < ks0 + ks_n; ++ks) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
                        const uint32_t b_dst = a_dst + A_BYTES;
                        mbar_expect_tx(full_bar(stage), STAGE_BYTES);
                        const int r0 = ks * BLOCK_K;
                        if (A_MN) {
#pragma unroll
                            for (int a = 0; a < BLOCK_M / ATOM_MN; ++a)
                                tma_load_3d(a_dst + a * ATOM_BYTES, ma, full_bar(stage), i0 + a * ATOM_MN, r0, it.s);
                        } else {
                            tma_load_3d(a_dst, ma, full_bar(stage), r0, i0, it.s);
                        }
                        if (B_MN) {
#pragma unroll
                            for (int a = 0; a < BLOCK_N / ATOM_MN; ++a)
                                tma_load_3d(b_dst + a * ATOM_BYTES, mb, full_bar(stage), j0 + a * ATOM_MN, r0, it.s);
                        } else {
                            tma_load_3d(b_dst, mb, full_bar(stage), r0, j0, it.s);
                        }
                        if (++stage == kStages) stage = 0, phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(A_MN, B_MN, BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int64_t L = blockIdx.x; L < n_items; L += gridDim.x, ++iter) {
                const int acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                int ks0, ks_n;
                chunk_range(p, decode_item(p, L).chunk, ks0, ks_n);
                const int total_steps = ks_n * p.passes;  // split-precision: 3 passes into the same accumulator
                for (int ks = 0; ks < total_steps; ++ks) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_src = smem_base + stage * STAGE_BYTES;
                    const uint32_t b_src = a_src + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_bf16(d_tmem, operand_desc<A_MN>(a_src, k), operand_desc<B_MN>(b_src, k), idesc,
                                  (ks > 0 || k > 0) ? 1u : 0u);
                    umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs retire
                    if (++stage == kStages) stage = 0, phase ^= 1u;
                }
                umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
            }
        }
    } else {
        // ===================== epilogue warps: TMEM -> regs -> swizzled smem -> TMA store =====================
        using OutT = typename std::conditional<OUT_F32, float, __nv_bfloat16>::type;
        constexpr int BOX_COLS = 128 / (int)sizeof(OutT);  // 32 fp32 or 64 bf16 columns = 128 B per row
        constexpr int LDS_PER_BOX = BOX_COLS / 32;
        constexpr int BOXES = BLOCK_N / BOX_COLS;
        const int q = warp & 3;         // TMEM lane quarter this warp may touch
        const int row = q * 32 + lane;  // row of the tile held by this thread
        const bool store_thread = (warp == 2 && lane == 0);
        int iter = 0;
        uint32_t box_count = 0;
        for (int64_t L = blockIdx.x; L < n_items; L += gridDim.x, ++iter) {
            const Item it = decode_item(p, L);
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int i0 = it.i_blk * BLOCK_M, j0 = it.j_blk * BLOCK_N;
            if (HAS_BIAS) {
                // all 128 epilogue threads need the same 256 bias values: fetch them once per tile, before the
                // accumulator wait (the previous tile's last named barrier already ordered the old reads)
                const int et = threadIdx.x - 64;  // 0..127
#pragma unroll
                for (int u = 0; u < BLOCK_N / (EPI_WARPS * 32); ++u) {
                    const int64_t jc = (int64_t)j0 + et + u * (EPI_WARPS * 32);
                    bias_gen[et + u * (EPI_WARPS * 32)] =
                        (jc < p.J && it.chunk == 0) ? __ldg(p.bias + (int64_t)it.s * p.J + jc) : 0.0f;
                }
            }
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            int n_boxes = BOXES;  // boxes that intersect the matrix (identical for all epilogue warps)
            if ((int64_t)j0 + BLOCK_N > p.J) n_boxes = (int)((p.J - j0 + BOX_COLS - 1) / BOX_COLS);
#pragma unroll 1
            for (int b = 0; b < n_boxes; ++b, ++box_count) {
                const uint32_t buf = box_count & 1u;
                uint32_t r[LDS_PER_BOX][32];
#pragma unroll
                for (int h = 0; h < LDS_PER_BOX; ++h) tmem_ld_32x32(t_acc + (uint32_t)(b * BOX_COLS + h * 32), r[h]);
                // the store issued two boxes ago read this buffer: it must have finished reading
                if (store_thread) tma_store_wait_read<1>();
                tmem_ld_wait();
                if (b == n_boxes - 1) {  // accumulator fully drained: hand the TMEM buffer back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
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
                        for (int ch = 0; ch < 8; ++ch)  // 8 x 16 B chunks, XOR-swizzled by (row % 8) like the TMA box
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
                fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA engine
                named_bar_sync<1, EPI_WARPS * 32>();
                if (store_thread) {
                    if (p.accumulate || p.r_chunks > 1)
                        tma_reduce_add_3d(&map_out, out_base + buf * OUT_BOX_BYTES, j0 + b * BOX_COLS, i0, it.s);
                    else tma_store_3d(&map_out, out_base + buf * OUT_BOX_BYTES, j0 + b * BOX_COLS, i0, it.s);
                    tma_store_commit();
                }
            }
        }
        if (store_thread) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <bool A_MN, bool B_MN, bool OUT_F32, bool HAS_BIAS>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, Params p, cudaStream_t st,
                  const CUtensorMap* ma_lo = nullptr, const CUtensorMap* mb_lo = nullptr) {
    auto kern = bayes_gemm_kernel<A_MN, B_MN, OUT_F32, HAS_BIAS>;
    BF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    p.passes = (ma_lo && mb_lo) ? 3 : 1;
    if (p.r_chunks < 1) p.r_chunks = 1, p.ks_per_chunk = p.k_steps;
    const int64_t n_items = p.S * p.r_chunks * p.i_tiles * p.j_tiles;
    const int64_t sms = bf_num_sms();
    const int grid = (int)(n_items < sms ? n_items : sms);
    kern<<<grid, kThreads, SMEM_BYTES, st>>>(ma, mb, mo, ma_lo ? *ma_lo : ma, mb_lo ? *mb_lo : mb, p);
    BF_LAUNCH_OK();
    return 0;
}

template <bool A_MN, bool B_MN>
static int launch_out(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const Params& p,
                      bool out_f32, cudaStream_t st, const CUtensorMap* ma_lo = nullptr,
                      const CUtensorMap* mb_lo = nullptr) {
    const bool bias = p.bias != nullptr;
    if (out_f32)
        return bias ? launch<A_MN, B_MN, true, true>(ma, mb, mo, p, st, ma_lo, mb_lo)
                    : launch<A_MN, B_MN, true, false>(ma, mb, mo, p, st, ma_lo, mb_lo);
    return bias ? launch<A_MN, B_MN, false, true>(ma, mb, mo, p, st, ma_lo, mb_lo)
                : launch<A_MN, B_MN, false, false>(ma, mb, mo, p, st, ma_lo, mb_lo);
}

}  // namespace tc

int bf_linear_fwd_bf16_2cta(const void*, const void*, const float*, void*, int64_t, int64_t, int64_t, int64_t, int32_t,
                            cudaStream_t);
int bf_linear_dgrad_bf16_2cta(const void*, const void*, void*, int64_t, int64_t, int64_t, int64_t, int32_t, int, cudaStream_t);

// CTA-pair (cta_group::2, 256 x 256 tiles) kernels of bf_gemm_tc2.cu when there are enough 256 x 256 tiles to give
// every SM pair one (measured: 3-10 % faster than the single-CTA kernel from there on, slower below);
// bf_set_option(BF_OPT_GEMM_2CTA, 0 / 2) forces the single-CTA / CTA-pair kernels (A/B measurements)
static bool use_cta_pairs(int64_t S, int64_t rows, int64_t cols) {
    const int mode = bf_option(BF_OPT_GEMM_2CTA);
    if (mode == 0 || rows < 256) return false;
    if (mode == 2) return true;
    const int64_t pair_tiles = S * ((rows + 255) / 256) * ((cols + 255) / 256);
    return pair_tiles >= bf_num_sms() / 2;
}

// y[s] = x[s] w[s]^T + bias[s]
int bf_linear_fwd_bf16(const void* x, const void* w, const float* bias, void* y, int64_t S, int64_t M, int64_t N,
                       int64_t K, int32_t y_dtype, cudaStream_t st) {
    using namespace tc;
    BF_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "bf16 path needs K % 8 == 0 and N % 8 == 0 (16 B TMA row strides)");
    if (use_cta_pairs(S, M, N)) return bf_linear_fwd_bf16_2cta(x, w, bias, y, S, M, N, K, y_dtype, st);
    CUtensorMap ma, mb, mo;
    int rc;
    if ((rc = encode_map(&ma, x, S, M, K, BLOCK_M))) return rc;
    if ((rc = encode_map(&mb, w, S, N, K, BLOCK_N))) return rc;
    if ((rc = encode_map(&mo, y, S, M, N, BLOCK_M, y_dtype == BF_F32))) return rc;
    Params p{};
    p.S = S, p.I = M, p.J = N, p.R = K;
    p.i_tiles = cdiv(M, BLOCK_M), p.j_tiles = cdiv(N, BLOCK_N), p.k_steps = cdiv(K, BLOCK_K);
    p.bias = bias;
    return launch_out<false, false>(ma, mb, mo, p, y_dtype == BF_F32, st);
}

// dx[s] = gy[s] w[s]
int bf_linear_dgrad_bf16(const void* gy, const void* w, void* dx, int64_t S, int64_t M, int64_t N, int64_t K,
                         int32_t dx_dtype, int accumulate, cudaStream_t st) {
    using namespace tc;
    BF_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "bf16 path needs K % 8 == 0 and N % 8 == 0");
    if (use_cta_pairs(S, M, K)) return bf_linear_dgrad_bf16_2cta(gy, w, dx, S, M, N, K, dx_dtype, accumulate, st);
    CUtensorMap ma, mb, mo;
    int rc;
    if ((rc = encode_map(&ma, gy, S, M, N, BLOCK_M))) return rc;  // A: K-major over r = N
    if ((rc = encode_map(&mb, w, S, N, K, BLOCK_K))) return rc;   // B: MN-major, rows = r = N, cols = K
    if ((rc = encode_map(&mo, dx, S, M, K, BLOCK_M, dx_dtype == BF_F32))) return rc;
    Params p{};
    p.S = S, p.I = M, p.J = K, p.R = N;
    p.i_tiles = cdiv(M, BLOCK_M), p.j_tiles = cdiv(K, BLOCK_N), p.k_steps = cdiv(N, BLOCK_K);
    p.accumulate = accumulate;
    return launch_out<false, true>(ma, mb, mo, p, dx_dtype == BF_F32, st);
}

// dw[s] = gy[s]^T x[s]   (unfused form: fp32 [S,N,K] to HBM)
int bf_linear_wgrad_bf16(const void* gy, const void* x, float* dw, int64_t S, int64_t M, int64_t N, int64_t K,
                         cudaStream_t st) {
    using namespace tc;
    BF_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "bf16 path needs K % 8 == 0 and N % 8 == 0");
    CUtensorMap ma, mb, mo;
    int rc;
    if ((rc = encode_map(&ma, gy, S, M, N, BLOCK_K))) return rc;  // A: MN-major, rows = r = M, cols = N
    if ((rc = encode_map(&mb, x, S, M, K, BLOCK_K))) return rc;   // B: MN-major, rows = r = M, cols = K
    if ((rc = encode_map(&mo, dw, S, N, K, BLOCK_M, true))) return rc;
    Params p{};
    p.S = S, p.I = N, p.J = K, p.R = M;
    p.i_tiles = cdiv(N, BLOCK_M), p.j_tiles = cdiv(K, BLOCK_N), p.k_steps = cdiv(M, BLOCK_K);
    return launch_out<true, true>(ma, mb, mo, p, true, st);
}

// ------------------------------------------------------------------ split-precision ("fp32x3") contractions
// fp32 operands travel as bf16 (hi, lo) pairs (bf_split_bf16x2); three tcgen05 passes per tile accumulate
// A_hi B_hi + A_hi B_lo + A_lo B_hi in fp32 (single-CTA 128 x 256 kernel, fp32 results).  Reference precision
// (bayeformers/nn/layers/linear.py:104 computes in fp32 with TF32 off) at tensor-core speed: the parity mode that
// the FFMA kernels of bf_gemm_simt.cu serve at 2 % of the tensor peak.
namespace {
__global__ void __launch_bounds__(256) split_bf16x2_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                                           __nv_bfloat16* __restrict__ lo, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * 256 * 4;
    for (int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n && (reinterpret_cast<uintptr_t>(src + i) & 15u) == 0 &&
            (reinterpret_cast<uintptr_t>(hi + i) & 7u) == 0 && (reinterpret_cast<uintptr_t>(lo + i) & 7u) == 0) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
            const float f[4] = {v.x, v.y, v.z, v.w};
            __nv_bfloat16 h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                h[j] = __float2bfloat16_rn(f[j]);
                l[j] = __float2bfloat16_rn(f[j] - __bfloat162float(h[j]));
            }
            *reinterpret_cast<uint2*>(hi + i) = *reinterpret_cast<const uint2*>(h);
            *reinterpret_cast<uint2*>(lo + i) = *reinterpret_cast<const uint2*>(l);
        } else {
            for (int j = 0; j < 4 && i + j < n; ++j) {
                const float f = src[i + j];
                const __nv_bfloat16 h = __float2bfloat16_rn(f);
                hi[i + j] = h;
                lo[i + j] = __float2bfloat16_rn(f - __bfloat162float(h));
            }
        }
    }
}
}  // namespace

extern "C" int bf_split_bf16x2(const float* src, void* hi, void* lo, int64_t n, void* stream) {
    BF_CHECK_ARG(n >= 0, "bad n");
    if (n == 0) return 0;
    BF_CHECK_ARG(src && hi && lo, "null pointer");
    const int64_t want = (n + 1023) / 1024, cap = (int64_t)bf_num_sms() * 16;
    split_bf16x2_kernel<<<(int)(want < cap ? want : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        src, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), n);
    BF_LAUNCH_OK();
    return 0;
}

#define BF_X3_CHECK()                                                                            \
    BF_CHECK_ARG(S >= 1 && M >= 1 && N >= 1 && K >= 1, "S, M, N, K must be >= 1");                 \
    BF_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "tensor-core path needs K % 8 == 0 and N % 8 == 0")

// at most kX3ChunkSteps k-steps (x 4 MMAs x 3 passes) are accumulated inside the tensor core; longer reductions are
// chunked and the chunk tiles added in global memory (Params::r_chunks), which needs a zero-filled output
constexpr int kX3ChunkSteps = 8;
static int x3_chunks(tc::Params& p, void* out, size_t out_bytes, cudaStream_t st) {
    p.ks_per_chunk = kX3ChunkSteps;
    p.r_chunks = tc::cdiv(p.k_steps, kX3ChunkSteps);
    if (p.r_chunks <= 1) {
        p.r_chunks = 1, p.ks_per_chunk = p.k_steps;
        return 0;
    }
    BF_CUDA_OK(cudaMemsetAsync(out, 0, out_bytes, st));
    return 0;
}

// y[s] = x[s] w[s]^T + bias[s]
extern "C" int bf_linear_fwd_x3(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                                const float* bias, float* y, int64_t S, int64_t M, int64_t N, int64_t K, void* stream) {
    using namespace tc;
    BF_CHECK_ARG(x_hi && x_lo && w_hi && w_lo && y, "null pointer");
    BF_X3_CHECK();
    CUtensorMap ma, mb, mo, mal, mbl;
    int rc;
    if ((rc = encode_map(&ma, x_hi, S, M, K, BLOCK_M))) return rc;
    if ((rc = encode_map(&mal, x_lo, S, M, K, BLOCK_M))) return rc;
    if ((rc = encode_map(&mb, w_hi, S, N, K, BLOCK_N))) return rc;
    if ((rc = encode_map(&mbl, w_lo, S, N, K, BLOCK_N))) return rc;
    if ((rc = encode_map(&mo, y, S, M, N, BLOCK_M, true))) return rc;
    Params p{};
    p.S = S, p.I = M, p.J = N, p.R = K;
    p.i_tiles = cdiv(M, BLOCK_M), p.j_tiles = cdiv(N, BLOCK_N), p.k_steps = cdiv(K, BLOCK_K);
    p.bias = bias;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if ((rc = x3_chunks(p, y, (size_t)S * M * N * sizeof(float), st))) return rc;
    return launch_out<false, false>(ma, mb, mo, p, true, st, &mal, &mbl);
}

// dx[s] = gy[s] w[s]
extern "C" int bf_linear_dgrad_x3(const void* gy_hi, const void* gy_lo, const void* w_hi, const void* w_lo, float* dx,
                                  int64_t S, int64_t M, int64_t N, int64_t K, void* stream) {
    using namespace tc;
    BF_CHECK_ARG(gy_hi && gy_lo && w_hi && w_lo && dx, "null pointer");
    BF_X3_CHECK();
    CUtensorMap ma, mb, mo, mal, mbl;
    int rc;
    if ((rc = encode_map(&ma, gy_hi, S, M, N, BLOCK_M))) return rc;
    if ((rc = encode_map(&mal, gy_lo, S, M, N, BLOCK_M))) return rc;
    if ((rc = encode_map(&mb, w_hi, S, N, K, BLOCK_K))) return rc;
    if ((rc = encode_map(&mbl, w_lo, S, N, K, BLOCK_K))) return rc;
    if ((rc = encode_map(&mo, dx, S, M, K, BLOCK_M, true))) return rc;
    Params p{};
    p.S = S, p.I = M, p.J = K, p.R = N;
    p.i_tiles = cdiv(M, BLOCK_M), p.j_tiles = cdiv(K, BLOCK_N), p.k_steps = cdiv(N, BLOCK_K);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if ((rc = x3_chunks(p, dx, (size_t)S * M * K * sizeof(float), st))) return rc;
    return launch_out<false, true>(ma, mb, mo, p, true, st, &mal, &mbl);
}

// dw[s] = gy[s]^T x[s]
extern "C" int bf_linear_wgrad_x3(const void* gy_hi, const void* gy_lo, const void* x_hi, const void* x_lo, float* dw,
                                  int64_t S, int64_t M, int64_t N, int64_t K, void* stream) {
    using namespace tc;
    BF_CHECK_ARG(gy_hi && gy_lo && x_hi && x_lo && dw, "null pointer");
    BF_X3_CHECK();
    CUtensorMap ma, mb, mo, mal, mbl;
    int rc;
    if ((rc = encode_map(&ma, gy_hi, S, M, N, BLOCK_K))) return rc;
    if ((rc = encode_map(&mal, gy_lo, S, M, N, BLOCK_K))) return rc;
    if ((rc = encode_map(&mb, x_hi, S, M, K, BLOCK_K))) return rc;
    if ((rc = encode_map(&mbl, x_lo, S, M, K, BLOCK_K))) return rc;
    if ((rc = encode_map(&mo, dw, S, N, K, BLOCK_M, true))) return rc;
    Params p{};
    p.S = S, p.I = N, p.J = K, p.R = M;
    p.i_tiles = cdiv(N, BLOCK_M), p.j_tiles = cdiv(K, BLOCK_N), p.k_steps = cdiv(M, BLOCK_K);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if ((rc = x3_chunks(p, dw, (size_t)S * N * K * sizeof(float), st))) return rc;
    return launch_out<true, true>(ma, mb, mo, p, true, st, &mal, &mbl);
}
