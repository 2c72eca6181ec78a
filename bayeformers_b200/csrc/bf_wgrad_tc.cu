// Fused variational weight gradient on tcgen05 (sm_100a).
//
// Replaces, for all S Monte-Carlo samples of one Bayesian Linear, what autograd
// does in the reference for  w = mu + softplus(rho) * eps ;  y = F.linear(x, w)
// (bayeformers/nn/parameters/gaussian.py:101, bayeformers/nn/layers/linear.py:104):
//
//     dW_s     = gy_s^T x_s                                  (tensor cores, never written to HBM)
//     grad_mu  = sum_s dW_s
//     grad_rho = sigmoid(rho) * sum_s dW_s o eps_s           (eps_s regenerated from the Philox counter)
//
// Structure: 128 x 128 output tiles, 5-stage TMA ring, tcgen05.mma into a
// double-buffered TMEM accumulator.  The other half of TMEM holds two running
// sums per tile (sum_s dW_s o eps_s and sum_s dW_s): after each sample's
// contraction the epilogue warps pull the accumulator (tcgen05.ld), generate
// the matching eps quads, FMA into the running sums and write them back with
// tcgen05.st -- no global traffic per sample.  Only when the last sample of a
// work item is done are the sums staged through shared memory and written
// (or read-modify-written) to grad_rho / grad_mu with coalesced 128 B rows.
//
// Small layers do not have enough tiles to fill 148 SMs, so a tile may be cut
// into several work items (groups of samples, then slices of the reduction).
// Items of one tile add their contribution to global memory in a FIXED order
// (turn counters), which keeps the result run-to-run deterministic without
// float atomics.
#include "bf_tc.cuh"

int bf_sample_kl_bwd_impl_kl_only(const float* mu, const float* rho, int32_t prior_kind, const float* prior_mu,
                                  const float* prior_rho, float pi, float sigma1, float sigma2, const float* g_logq,
                                  const float* g_logp, int64_t n, int32_t S, uint64_t seed, uint32_t step,
                                  uint32_t tensor_id, const float* eps_in, float* grad_mu, float* grad_rho,
                                  cudaStream_t st);

namespace wg {
using namespace tc;

constexpr int BM = 128;
constexpr int EPI_COLS = 32, EPI_WARPS = 8;  // 2 warps per TMEM lane quarter, each owning half of the column chunks
constexpr int kThreads = 32 * (2 + EPI_WARPS);
constexpr int TMEM_COLS = 512;

// Two tile shapes (TMEM is 512 columns in both):
//   BN = 256 (mu frozen, the MOPED fine-tuning case): [0,256) accumulator | [256,512) sum dW*eps.
//            128x256 tiles have 1.37x the arithmetic intensity of 128x128 ones, which is what the
//            L2->SM operand stream (measured ~11-14 TB/s with everything in flight) needs.
//   BN = 128 (mu trainable): [0,128) acc0 | [128,256) acc1 | [256,384) sum dW*eps | [384,512) sum dW.
template <int BN>
struct Cfg {
    static constexpr int NACC = BN == 128 ? 2 : 1;
    static constexpr int A_BYTES = BM * BLOCK_K * 2, B_BYTES = BN * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int kStages = BN == 128 ? 6 : 4;
    static constexpr int T_SUM_RHO = 256, T_SUM_MU = 384;
    static constexpr int CHUNKS_PER_WARP = (BN / EPI_COLS) / 2;
    static constexpr int SMEM_BYTES = 1024 + kStages * STAGE_BYTES + 256;
};

struct Params {
    int64_t S, I, J, R;  // samples, rows of W (N), cols of W (K), reduction length (M)
    int i_tiles, j_tiles, k_steps, groups, splits;
    const float* rho;
    const float* eps_in;  // [S][I][J] injected eps or null
    float* grad_mu;       // null when mu is frozen
    float* grad_rho;
    int* turn;            // [i_tiles*j_tiles], zero-initialised, self-resetting
    int accumulate;
    uint32_t k0, k1, step, tensor_id;
    const uint32_t* step_ptr;  // optional device-resident offset added to `step`
};

struct Item {
    int tile, turn, turns, i_blk, j_blk, s_begin, s_end, k_begin, k_end;
};
__device__ __forceinline__ Item decode_item(const Params& p, int64_t L) {
    Item it;
    it.turns = p.groups * p.splits;
    it.turn = (int)(L % it.turns);
    it.tile = (int)(L / it.turns);
    const int g = it.turn / p.splits, sp = it.turn % p.splits;
    it.j_blk = it.tile % p.j_tiles;
    it.i_blk = it.tile / p.j_tiles;
    it.s_begin = (int)((p.S * g) / p.groups);
    it.s_end = (int)((p.S * (g + 1)) / p.groups);
    it.k_begin = (int)(((int64_t)p.k_steps * sp) / p.splits);
    it.k_end = (int)(((int64_t)p.k_steps * (sp + 1)) / p.splits);
    return it;
}

template <int BN, bool HAS_EPS, bool WITH_MU>
__global__ void __launch_bounds__(kThreads, 1)
    bayes_wgrad_kernel(const __grid_constant__ CUtensorMap map_gy, const __grid_constant__ CUtensorMap map_x,
                       const __grid_constant__ Params p) {
    using C = Cfg<BN>;
    static_assert(!(WITH_MU && BN != 128), "the mu sum only fits next to 128-column accumulators");
    constexpr int kStages = C::kStages, NACC = C::NACC;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + kStages * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    volatile uint32_t* const tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * C::STAGE_BYTES + 8 * (2 * kStages + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_gy);
        tma_prefetch_desc(&map_x);
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
    const int64_t n_items = (int64_t)p.i_tiles * p.j_tiles * p.groups * p.splits;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t L = blockIdx.x; L < n_items; L += gridDim.x) {
                const Item it = decode_item(p, L);
                const int i0 = it.i_blk * BM, j0 = it.j_blk * BN;
                for (int s = it.s_begin; s < it.s_end; ++s) {
                    for (int ks = it.k_begin; ks < it.k_end; ++ks) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES;
                        const uint32_t b_dst = a_dst + C::A_BYTES;
                        mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
                        const int r0 = ks * BLOCK_K;
#pragma unroll
                        for (int a = 0; a < BM / ATOM_MN; ++a)  // gy[s][m][n]: MN-major, rows = reduction m
                            tma_load_3d(a_dst + a * ATOM_BYTES, &map_gy, full_bar(stage), i0 + a * ATOM_MN, r0, s);
#pragma unroll
                        for (int a = 0; a < BN / ATOM_MN; ++a)  // x[s][m][k]
                            tma_load_3d(b_dst + a * ATOM_BYTES, &map_x, full_bar(stage), j0 + a * ATOM_MN, r0, s);
                        if (++stage == kStages) stage = 0, phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(true, true, BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;  // counts (item, sample) pairs
            for (int64_t L = blockIdx.x; L < n_items; L += gridDim.x) {
                const Item it = decode_item(p, L);
                for (int s = it.s_begin; s < it.s_end; ++s, ++iter) {
                    const int acc = iter % NACC;
                    const uint32_t acc_phase = (iter / NACC) & 1;
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                    for (int ks = it.k_begin; ks < it.k_end; ++ks) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t a_src = smem_base + stage * C::STAGE_BYTES;
                        const uint32_t b_src = a_src + C::A_BYTES;
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_bf16(d_tmem, operand_desc<true>(a_src, k), operand_desc<true>(b_src, k), idesc,
                                      (ks > it.k_begin || k > 0) ? 1u : 0u);
                        umma_commit(empty_bar(stage));
                        if (++stage == kStages) stage = 0, phase ^= 1u;
                    }
                    umma_commit(tfull_bar(acc));
                }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        // warps 2..9: TMEM lane quarter = warp % 4 (hardware rule); the two warps of a quarter split the
        // column chunks, and two warps per scheduler hide each other's dependency stalls
        const uint32_t step = p.step + (p.step_ptr ? __ldg(p.step_ptr) : 0u);
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        int iter = 0;
        for (int64_t L = blockIdx.x; L < n_items; L += gridDim.x) {
            const Item it = decode_item(p, L);
            const int64_t j_base = (int64_t)it.j_blk * BN;
            const int64_t my_row = (int64_t)it.i_blk * BM + q * 32 + lane;
            const bool row_ok = my_row < p.I;
            int n_chunks = BN / EPI_COLS;  // warp-uniform
            if (j_base + BN > p.J) n_chunks = (int)((p.J - j_base + EPI_COLS - 1) / EPI_COLS);
            const int c_lo = half * C::CHUNKS_PER_WARP;
            const int c_hi = min(n_chunks, c_lo + C::CHUNKS_PER_WARP);

            // ---- per sample: sums (TMEM) += accumulator (TMEM) o eps (registers) ----
            for (int s = it.s_begin; s < it.s_end; ++s, ++iter) {
                const int acc = iter % NACC;
                const uint32_t acc_phase = (iter / NACC) & 1;
                const bool first = (s == it.s_begin);
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                {
#pragma unroll 1
                    for (int c = c_lo; c < c_hi; ++c) {
                        uint32_t a[32], sr[32], sm[32];
                        tmem_ld_32x32(lane_base + (uint32_t)(acc * BN + c * EPI_COLS), a);
                        if (!first) {
                            tmem_ld_32x32(lane_base + (uint32_t)(C::T_SUM_RHO + c * EPI_COLS), sr);
                            if (WITH_MU) tmem_ld_32x32(lane_base + (uint32_t)(C::T_SUM_MU + c * EPI_COLS), sm);
                        }
                        float e[32];
                        const int64_t flat = my_row * p.J + j_base + c * EPI_COLS;
                        if (row_ok) {
#pragma unroll
                            for (int t = 0; t < 8; ++t) {
                                float4 v;
                                if (HAS_EPS) {
                                    v = (j_base + c * EPI_COLS + 4 * t + 4 <= p.J)
                                            ? __ldg(reinterpret_cast<const float4*>(p.eps_in + (int64_t)s * p.I * p.J +
                                                                                     flat + 4 * t))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                                } else {
                                    v = bf_eps_quad((uint32_t)((flat >> 2) + t), (uint32_t)s, p.tensor_id, step, p.k0, p.k1);
                                }
                                e[4 * t] = v.x, e[4 * t + 1] = v.y, e[4 * t + 2] = v.z, e[4 * t + 3] = v.w;
                            }
                        } else {
#pragma unroll
                            for (int t = 0; t < 32; ++t) e[t] = 0.0f;
                        }
                        tmem_ld_wait();
#pragma unroll
                        for (int t = 0; t < 32; ++t) {
                            const float av = __uint_as_float(a[t]);
                            const float prev = first ? 0.0f : __uint_as_float(sr[t]);
                            sr[t] = __float_as_uint(fmaf(av, e[t], prev));
                            if (WITH_MU) sm[t] = __float_as_uint(first ? av : __uint_as_float(sm[t]) + av);
                        }
                        tmem_st_32x32(lane_base + (uint32_t)(C::T_SUM_RHO + c * EPI_COLS), sr);
                        if (WITH_MU) tmem_st_32x32(lane_base + (uint32_t)(C::T_SUM_MU + c * EPI_COLS), sm);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));  // accumulator drained
                }
            }

            // ---- once per item: -> global gradients, items of one tile in fixed turn order ----
            if (it.turns > 1) {
                if (lane == 0) {
                    const volatile int* t = p.turn + it.tile;
                    while (*t != it.turn) __nanosleep(64);
                    __threadfence();
                }
                __syncwarp();
            }
            const bool add = (it.turn > 0) || p.accumulate;
#pragma unroll 1
            for (int kind = 0; kind < (WITH_MU ? 2 : 1); ++kind) {  // 0: grad_rho (x sigmoid(rho)), 1: grad_mu
                float* const dst = kind == 0 ? p.grad_rho : p.grad_mu;
                const int t_off = kind == 0 ? C::T_SUM_RHO : C::T_SUM_MU;
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    uint32_t r[32];
                    const int64_t jc = j_base + c * EPI_COLS;
                    const int64_t flat = my_row * p.J + jc;
                    tmem_ld_32x32(lane_base + (uint32_t)(t_off + c * EPI_COLS), r);
                    // thread = one row of W, 32 consecutive columns = one full 128 B line per thread.
                    // All loads of the chunk are issued before any store (the compiler cannot hoist the
                    // read-modify-write loads above stores to the same array on its own).
                    float4 scale[8], old[8];
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const bool in = row_ok && (jc + 4 * v + 4 <= p.J);  // J % 4 == 0
                        scale[v] = (in && kind == 0) ? __ldg(reinterpret_cast<const float4*>(p.rho + flat + 4 * v))
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
                        old[v] = (in && add) ? __ldcg(reinterpret_cast<const float4*>(dst + flat + 4 * v))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    tmem_ld_wait();  // warp-collective: must stay outside any lane-divergent branch
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        if (row_ok && jc + 4 * v + 4 <= p.J) {
                            float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
                            if (kind == 0) {
                                g.x = bf_softplus_grad(scale[v].x), g.y = bf_softplus_grad(scale[v].y);
                                g.z = bf_softplus_grad(scale[v].z), g.w = bf_softplus_grad(scale[v].w);
                            }
                            float4 o;
                            o.x = fmaf(__uint_as_float(r[4 * v + 0]), g.x, old[v].x);
                            o.y = fmaf(__uint_as_float(r[4 * v + 1]), g.y, old[v].y);
                            o.z = fmaf(__uint_as_float(r[4 * v + 2]), g.z, old[v].z);
                            o.w = fmaf(__uint_as_float(r[4 * v + 3]), g.w, old[v].w);
                            __stcg(reinterpret_cast<float4*>(dst + flat + 4 * v), o);
                        }
                    }
                }
            }
            if (it.turns > 1) {
                __threadfence();
                named_bar_sync<1, EPI_WARPS * 32>();
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

template <int BN, bool HAS_EPS, bool WITH_MU>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const Params& p, cudaStream_t st) {
    auto kern = bayes_wgrad_kernel<BN, HAS_EPS, WITH_MU>;
    BF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
    const int64_t n_items = (int64_t)p.i_tiles * p.j_tiles * p.groups * p.splits;
    const int64_t sms = bf_num_sms();
    kern<<<(int)(n_items < sms ? n_items : sms), kThreads, Cfg<BN>::SMEM_BYTES, st>>>(ma, mb, p);
    BF_LAUNCH_OK();
    return 0;
}

}  // namespace wg

int64_t bf_wgrad_fused_workspace_ints(int64_t N, int64_t K) {
    return (int64_t)tc::cdiv(N, wg::BM) * tc::cdiv(K, 128);  // enough for either tile width
}

int bf_linear_wgrad_fused_bf16(const void* gy, const void* x, int64_t S, int64_t M, int64_t N, int64_t K,
                               const float* mu, const float* rho, int32_t prior_kind, const float* prior_mu,
                               const float* prior_rho, float pi, float sigma1, float sigma2, const float* g_logq,
                               const float* g_logp, uint64_t seed, uint32_t step, uint32_t tensor_id,
                               const float* eps_in, float* grad_mu, float* grad_rho, int32_t accumulate, int* turn_ws,
                               cudaStream_t st) {
    using namespace wg;
    BF_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "bf16 path needs K % 8 == 0 and N % 8 == 0");
    BF_CHECK_ARG(turn_ws != nullptr, "turn workspace missing");
    CUtensorMap ma, mb;
    int rc;
    if ((rc = tc::encode_map(&ma, gy, S, M, N, tc::BLOCK_K))) return rc;
    if ((rc = tc::encode_map(&mb, x, S, M, K, tc::BLOCK_K))) return rc;
    const bool eps = eps_in != nullptr, with_mu = grad_mu != nullptr;
    const int bn = with_mu ? 128 : 256;
    Params p{};
    p.S = S, p.I = N, p.J = K, p.R = M;
    p.i_tiles = tc::cdiv(N, BM), p.j_tiles = tc::cdiv(K, bn), p.k_steps = tc::cdiv(M, tc::BLOCK_K);
    // fill the persistent grid: first cut a tile by groups of samples (each (element, sample) eps is still
    // generated exactly once), then by slices of the reduction (eps regenerated per slice)
    const int64_t tiles = (int64_t)p.i_tiles * p.j_tiles;
    const int64_t sms = bf_num_sms();
    int groups = 1, splits = 1;
    while (tiles * groups * 2 <= sms && groups * 2 <= S) groups *= 2;
    while (tiles * groups * splits * 2 <= sms && p.k_steps / (splits * 2) >= 16) splits *= 2;
    p.groups = groups, p.splits = splits;
    p.rho = rho, p.eps_in = eps_in, p.grad_mu = grad_mu, p.grad_rho = grad_rho, p.turn = turn_ws;
    p.accumulate = accumulate;
    p.k0 = (uint32_t)(seed & 0xffffffffu), p.k1 = (uint32_t)(seed >> 32), p.step = step, p.tensor_id = tensor_id;
    p.step_ptr = bf_step_counter();
    if (with_mu)
        rc = eps ? launch<128, true, true>(ma, mb, p, st) : launch<128, false, true>(ma, mb, p, st);
    else
        rc = eps ? launch<256, true, false>(ma, mb, p, st) : launch<256, false, false>(ma, mb, p, st);
    if (rc) return rc;
    // KL terms do not involve gy: they are one elementwise pass added on top (eps regenerated once more)
    if (g_logq != nullptr || g_logp != nullptr)
        return bf_sample_kl_bwd_impl_kl_only(mu, rho, prior_kind, prior_mu, prior_rho, pi, sigma1, sigma2, g_logq, g_logp,
                                             N * K, (int32_t)S, seed, step, tensor_id, eps_in, grad_mu, grad_rho, st);
    return 0;
}
