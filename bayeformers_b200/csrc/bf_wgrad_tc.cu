// Fused variational weight gradient on tcgen05 (sm_100a).
//
// Replaces, for all S Monte-Carlo samples of one Bayesian Linear, what autograd
// does in the reference for  w = mu + softplus(rho) * eps ;  y = F.linear(x, w)
// (bayeformers/nn/parameters/gaussian.py:101, bayeformers/nn/layers/linear.py:104):
//
//     dW_s     = gy_s^T x_s                                  (tensor cores; the raw dW_s never reaches HBM
//                                                             unless mu is trainable)
//     grad_mu  = sum_s dW_s                      (+ KL terms)
//     grad_rho = sigmoid(rho) * sum_s dW_s o eps_s  (+ KL terms; eps_s regenerated from the Philox counter)
//
// Two launches:
//
//  1. bayes_wgrad_kernel -- persistent, warp-specialised (1 TMA warp, 1 MMA warp,
//     8 epilogue warps), 128 x 256 output tiles, 4-stage TMA ring, tcgen05.mma into
//     a DOUBLE-BUFFERED TMEM accumulator (2 x 256 columns).  A work item is
//     (sample s, reduction slice sp, tile): small layers (768 x 768 has only 18
//     tiles) are cut along the reduction M = B*T so that every SM has work.  The
//     epilogue of item i overlaps the MMAs of item i+1: it pulls the accumulator
//     (tcgen05.ld), regenerates the matching eps quads with Philox, multiplies,
//     stages the 128 x 32 fp32 box in swizzled shared memory and TMA-stores it to
//     the partial buffer P[s*splits+sp][N][K].
//  2. wgrad_reduce_kernel -- one coalesced pass: grad_rho = sigmoid(rho) *
//     (sum_t P[t] + KL terms), in the FIXED order t = 0..T-1, so the result is
//     run-to-run deterministic (no float atomics, no turn-taking between CTAs).
//     The partials are written and re-read back to back; for BERT-sized layers
//     they are L2-resident (<= 75 MB against 126 MB).
#include <cstdlib>

#include "bf_tc.cuh"

namespace wg {
using namespace tc;

constexpr int BM = 128, BN = 256;
constexpr int A_BYTES = BM * BLOCK_K * 2;
constexpr int BOX_COLS = 32;                 // fp32 columns of one TMA-store box (128 B rows)
constexpr int BOX_BYTES = BM * 128;          // 16 KiB
constexpr int BOXES = BN / BOX_COLS;         // 8
constexpr int EPI_GROUPS = 2, EPI_WARPS = 4 * EPI_GROUPS;  // group g owns boxes b with b % 2 == g
constexpr int kThreads = 32 * (2 + EPI_WARPS);
constexpr int TMEM_COLS = 2 * BN;

// PAIR: two CTAs of a cluster share one 256-row MMA (cta_group::2, see bf_gemm_tc2.cu): each loads its 128 rows of
// gy^T and only 128 of the 256 x columns, 32 KiB per k-step instead of 48 KiB, so the ring gets 6 stages.
template <bool WITH_MU, bool PAIR>
struct Cfg {
    static constexpr int LOAD_N = PAIR ? BN / 2 : BN;  // x columns this CTA loads
    static constexpr int B_BYTES = LOAD_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BUFS = WITH_MU ? 2 : 1;  // staging boxes per epilogue group
    static constexpr int kStages = PAIR ? (WITH_MU ? 5 : 6) : (WITH_MU ? 3 : 4);
    static constexpr int SMEM_BYTES = 1024 + kStages * STAGE_BYTES + EPI_GROUPS * BUFS * BOX_BYTES + 256;
};

struct Params {
    int64_t S, I, J, R;  // samples, rows of W (N), cols of W (K), reduction length (M)
    int i_tiles, j_tiles, k_steps, splits;
    const float* eps_in;  // [S][I][J] injected eps or null
    uint32_t k0, k1, step, tensor_id;
    const uint32_t* step_ptr;  // optional device-resident offset added to `step`
};

struct Item {
    int s, sp, i_blk, j_blk, k_begin, k_end;
};
__device__ __forceinline__ Item decode_item(const Params& p, int64_t L) {
    Item it;
    it.j_blk = (int)(L % p.j_tiles);
    int64_t q = L / p.j_tiles;
    it.i_blk = (int)(q % p.i_tiles);
    q /= p.i_tiles;
    it.sp = (int)(q % p.splits);
    it.s = (int)(q / p.splits);
    it.k_begin = (int)(((int64_t)p.k_steps * it.sp) / p.splits);
    it.k_end = (int)(((int64_t)p.k_steps * (it.sp + 1)) / p.splits);
    return it;
}

template <bool HAS_EPS, bool WITH_MU, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
    bayes_wgrad_kernel(const __grid_constant__ CUtensorMap map_gy, const __grid_constant__ CUtensorMap map_x,
                       const __grid_constant__ CUtensorMap map_prho, const __grid_constant__ CUtensorMap map_pmu,
                       const __grid_constant__ Params p) {
    using C = Cfg<WITH_MU, PAIR>;
    constexpr int kStages = C::kStages, STAGE_BYTES = C::STAGE_BYTES;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs)
    const bool leader = rank == 0;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t out_base = smem_base + kStages * STAGE_BYTES;
    uint8_t* const out_gen = smem_gen + kStages * STAGE_BYTES;
    constexpr int OUT_BYTES = EPI_GROUPS * C::BUFS * BOX_BYTES;
    const uint32_t bar_base = out_base + OUT_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
    volatile uint32_t* const tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(out_gen + OUT_BYTES + 8 * (2 * kStages + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_gy);
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_prho);
        if (WITH_MU) tma_prefetch_desc(&map_pmu);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), (PAIR ? 2 : 1) * EPI_WARPS);  // PAIR: leader's copy collects both CTAs
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) {
        if (PAIR) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
        else tmem_alloc(tmem_slot, TMEM_COLS);
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    const int64_t n_items = p.S * p.splits * p.i_tiles * p.j_tiles;
    const int64_t worker = PAIR ? (blockIdx.x >> 1) : blockIdx.x, n_workers = PAIR ? (gridDim.x >> 1) : gridDim.x;

    if (warp == 0) {
        // ===================== TMA producer (every CTA: its own halves) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t L = worker; L < n_items; L += n_workers) {
                const Item it = decode_item(p, L);
                const int i0 = it.i_blk * (PAIR ? 2 * BM : BM) + (int)rank * BM;
                const int j0 = it.j_blk * BN + (int)rank * C::LOAD_N;
                for (int ks = it.k_begin; ks < it.k_end; ++ks) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
                    const uint32_t b_dst = a_dst + A_BYTES;
                    if (leader) mbar_expect_tx(full_bar(stage), (PAIR ? 2 : 1) * STAGE_BYTES);
                    const int r0 = ks * BLOCK_K;
#pragma unroll
                    for (int a = 0; a < BM / ATOM_MN; ++a) {  // gy[s][m][n]: MN-major, rows = reduction m
                        if (PAIR) tma_load_3d_2sm(a_dst + a * ATOM_BYTES, &map_gy, full_bar(stage), i0 + a * ATOM_MN, r0, it.s);
                        else tma_load_3d(a_dst + a * ATOM_BYTES, &map_gy, full_bar(stage), i0 + a * ATOM_MN, r0, it.s);
                    }
#pragma unroll
                    for (int a = 0; a < C::LOAD_N / ATOM_MN; ++a) {  // x[s][m][k]
                        if (PAIR) tma_load_3d_2sm(b_dst + a * ATOM_BYTES, &map_x, full_bar(stage), j0 + a * ATOM_MN, r0, it.s);
                        else tma_load_3d(b_dst + a * ATOM_BYTES, &map_x, full_bar(stage), j0 + a * ATOM_MN, r0, it.s);
                    }
                    if (++stage == kStages) stage = 0, phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA, one thread) =====================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc(true, true, PAIR ? 2 * BM : BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int64_t L = worker; L < n_items; L += n_workers, ++iter) {
                const Item it = decode_item(p, L);
                const int acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int ks = it.k_begin; ks < it.k_end; ++ks) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_src = smem_base + stage * STAGE_BYTES;
                    const uint32_t b_src = a_src + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint32_t accum = (ks > it.k_begin || k > 0) ? 1u : 0u;
                        if (PAIR) umma_bf16_2sm(d_tmem, operand_desc<true>(a_src, k), operand_desc<true>(b_src, k), idesc, accum);
                        else umma_bf16(d_tmem, operand_desc<true>(a_src, k), operand_desc<true>(b_src, k), idesc, accum);
                    }
                    if (PAIR) umma_commit_2sm(empty_bar(stage));
                    else umma_commit(empty_bar(stage));
                    if (++stage == kStages) stage = 0, phase ^= 1u;
                }
                if (PAIR) umma_commit_2sm(tfull_bar(acc));
                else umma_commit(tfull_bar(acc));
            }
        }
    } else {
        // ===================== epilogue warps =====================
        // warps 2..9: TMEM lane quarter = warp % 4 (hardware rule); epilogue group g = (warp-2)/4 takes the
        // boxes b with b % 2 == g, so the two warps that share a quarter split a tile's columns evenly.
        const uint32_t step = p.step + (p.step_ptr ? __ldg(p.step_ptr) : 0u);
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const bool store_thread = ((warp - 2) & 3) == 0 && lane == 0;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t my_out = out_base + grp * C::BUFS * BOX_BYTES;
        uint8_t* const my_out_gen = out_gen + grp * C::BUFS * BOX_BYTES;
        int iter = 0;
        for (int64_t L = worker; L < n_items; L += n_workers, ++iter) {
            const Item it = decode_item(p, L);
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int i0 = it.i_blk * (PAIR ? 2 * BM : BM) + (int)rank * BM, j0 = it.j_blk * BN;
            const int64_t g_row = (int64_t)i0 + row;
            const bool row_ok = g_row < p.I;
            int n_boxes = BOXES;  // boxes that intersect the matrix
            if ((int64_t)j0 + BN > p.J) n_boxes = (int)((p.J - j0 + BOX_COLS - 1) / BOX_COLS);
            const int out_z = it.s * p.splits + it.sp;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            int last_b = -1;  // last box of this group
            for (int b = grp; b < n_boxes; b += EPI_GROUPS) last_b = b;
            if (last_b < 0) {  // nothing to read: hand the accumulator straight back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) mbar_arrive_leader(tempty_bar(acc));
                    else mbar_arrive(tempty_bar(acc));
                }
            }
#pragma unroll 1
            for (int b = grp; b < n_boxes; b += EPI_GROUPS) {
                uint32_t r[32];
                tmem_ld_32x32(lane_base + (uint32_t)(acc * BN + b * BOX_COLS), r);
                // eps of this thread's 32 elements (row g_row, columns jc .. jc+31) while the TMEM load is in flight
                const int64_t jc = (int64_t)j0 + b * BOX_COLS;
                const int64_t flat = g_row * p.J + jc;
                float e[32];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    float4 v;
                    if (HAS_EPS) {
                        v = (row_ok && jc + 4 * t + 4 <= p.J)
                                ? __ldg(reinterpret_cast<const float4*>(p.eps_in + (int64_t)it.s * p.I * p.J + flat + 4 * t))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
                    } else {
                        v = bf_eps_quad((uint32_t)((flat >> 2) + t), (uint32_t)it.s, p.tensor_id, step, p.k0, p.k1);
                    }
                    e[4 * t] = v.x, e[4 * t + 1] = v.y, e[4 * t + 2] = v.z, e[4 * t + 3] = v.w;
                }
                tmem_ld_wait();
                if (b == last_b) {  // accumulator drained by this warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (PAIR) mbar_arrive_leader(tempty_bar(acc));
                        else mbar_arrive(tempty_bar(acc));
                    }
                }
                if (store_thread) tma_store_wait_read<0>();  // the previous store of this group has read its buffer
                named_bar_sync_dyn(1 + grp, 128);
                uint8_t* const my_row = my_out_gen + row * 128;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)  // 8 x 16 B chunks, XOR-swizzled by (row % 8) like the TMA box
                    *reinterpret_cast<float4*>(my_row + ((ch ^ (row & 7)) << 4)) =
                        make_float4(__uint_as_float(r[4 * ch]) * e[4 * ch], __uint_as_float(r[4 * ch + 1]) * e[4 * ch + 1],
                                    __uint_as_float(r[4 * ch + 2]) * e[4 * ch + 2],
                                    __uint_as_float(r[4 * ch + 3]) * e[4 * ch + 3]);
                if (WITH_MU) {
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch)
                        *reinterpret_cast<float4*>(my_row + BOX_BYTES + ((ch ^ (row & 7)) << 4)) =
                            make_float4(__uint_as_float(r[4 * ch]), __uint_as_float(r[4 * ch + 1]),
                                        __uint_as_float(r[4 * ch + 2]), __uint_as_float(r[4 * ch + 3]));
                }
                fence_proxy_async();
                named_bar_sync_dyn(1 + grp, 128);
                if (store_thread) {
                    tma_store_3d(&map_prho, my_out, (int)jc, i0, out_z);
                    if (WITH_MU) tma_store_3d(&map_pmu, my_out + BOX_BYTES, (int)jc, i0, out_z);
                    tma_store_commit();
                }
            }
        }
        if (store_thread) tma_store_wait_all();
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all();  // the peer's smem / TMEM must outlive the leader's last MMA
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <bool HAS_EPS, bool WITH_MU, bool PAIR>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mr, const CUtensorMap& mm,
                  const Params& p, cudaStream_t st) {
    auto kern = bayes_wgrad_kernel<HAS_EPS, WITH_MU, PAIR>;
    constexpr int smem = Cfg<WITH_MU, PAIR>::SMEM_BYTES;
    BF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t n_items = p.S * p.splits * p.i_tiles * p.j_tiles;
    const int64_t workers = PAIR ? bf_num_sms() / 2 : bf_num_sms();
    const int64_t use = n_items < workers ? n_items : workers;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(PAIR ? 2 * use : use));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    BF_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ma, mb, mr, mm, p));
    return 0;
}

// ---- how many slices of the reduction: fill the persistent grid, keep items long enough that the
// Philox epilogue (~10 us per 128x256 tile) stays hidden behind the next item's MMAs (0.22 us per k-step at peak)
static int choose_splits(int64_t S, int64_t tiles, int k_steps, int64_t workers, double* cost_out) {
    const int64_t base = S * tiles;
    const double fixed = 14.0;  // per-item overhead in k-step units (pipeline fill + exposed part of the epilogue)
    int best = 1;
    double best_cost = 1e30;
    for (int sp = 1; sp <= 16; ++sp) {
        if (sp > 1 && k_steps / sp < 24) break;
        const int64_t waves = (base * sp + workers - 1) / workers;
        const double cost = (double)waves * ((double)((k_steps + sp - 1) / sp) + fixed) + 0.5 * sp;
        if (cost < best_cost - 1e-9) best_cost = cost, best = sp;
    }
    if (cost_out) *cost_out = best_cost;
    return best;
}

// tile shape + slicing of one call: single CTAs on 128 x 256 tiles, or CTA pairs on 256 x 256 tiles (same time per
// item, a third less operand traffic) when that does not cost waves.  bf_set_option(BF_OPT_WGRAD_2CTA, 0 / 2) forces
// single / pair.
struct Plan {
    bool pair;
    int i_tiles, j_tiles, k_steps, splits;
};
static Plan make_plan(int64_t S, int64_t M, int64_t N, int64_t K) {
    const int mode = bf_option(BF_OPT_WGRAD_2CTA);
    Plan a{}, b{};
    double ca = 0, cb = 0;
    const int ks = tc::cdiv(M, tc::BLOCK_K);
    a.pair = false, a.i_tiles = tc::cdiv(N, BM), a.j_tiles = tc::cdiv(K, BN), a.k_steps = ks;
    a.splits = choose_splits(S, (int64_t)a.i_tiles * a.j_tiles, ks, bf_num_sms(), &ca);
    b.pair = true, b.i_tiles = tc::cdiv(N, 2 * BM), b.j_tiles = tc::cdiv(K, BN), b.k_steps = ks;
    b.splits = choose_splits(S, (int64_t)b.i_tiles * b.j_tiles, ks, bf_num_sms() / 2, &cb);
    if (mode == 0 || N < 2 * BM) return a;
    if (mode == 2) return b;
    return cb <= ca * 1.02 ? b : a;
}

constexpr int kRedThreads = 256;

struct ReduceParams {
    const float* part_rho;  // [T][n]
    const float* part_mu;   // [T][n] or null
    int T;
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
    int S, accumulate;
    uint32_t k0, k1, step, tensor_id;
    const uint32_t* step_ptr;
    float prior_const_ipv;
    BfMixture mix;
};

__device__ __forceinline__ float4 ld4s(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

// grad_rho = sigmoid(rho) * (sum_t P_rho[t] + sum_s KL_s),  grad_mu = sum_t P_mu[t] + sum_s g_logp[s] dlogp/dw(w_s)
// n % 4 == 0 (N*K with K % 8 == 0), all pointers 16 B aligned (torch allocations)
template <int PRIOR, bool KL>
__global__ void __launch_bounds__(kRedThreads) wgrad_reduce_kernel(const ReduceParams p) {
    const uint32_t step = p.step + (p.step_ptr ? __ldg(p.step_ptr) : 0u);
    const int64_t nquad = p.n >> 2;
    const int64_t stride = (int64_t)gridDim.x * kRedThreads;
    for (int64_t q = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; q < nquad; q += stride) {
        const int64_t i0 = q << 2;
        float4 ar = make_float4(0.f, 0.f, 0.f, 0.f), am = ar;
        {
            const float* src = p.part_rho + i0;
            int t = 0;
            for (; t + 4 <= p.T; t += 4) {  // 4 independent loads in flight, summed in fixed order
                const float4 a = ld4s(src + (int64_t)t * p.n), b = ld4s(src + (int64_t)(t + 1) * p.n);
                const float4 c = ld4s(src + (int64_t)(t + 2) * p.n), d = ld4s(src + (int64_t)(t + 3) * p.n);
                ar.x += a.x, ar.y += a.y, ar.z += a.z, ar.w += a.w;
                ar.x += b.x, ar.y += b.y, ar.z += b.z, ar.w += b.w;
                ar.x += c.x, ar.y += c.y, ar.z += c.z, ar.w += c.w;
                ar.x += d.x, ar.y += d.y, ar.z += d.z, ar.w += d.w;
            }
            for (; t < p.T; ++t) {
                const float4 a = ld4s(src + (int64_t)t * p.n);
                ar.x += a.x, ar.y += a.y, ar.z += a.z, ar.w += a.w;
            }
        }
        if (p.grad_mu != nullptr && p.part_mu != nullptr) {
            const float* src = p.part_mu + i0;
            for (int t = 0; t < p.T; ++t) {
                const float4 a = ld4s(src + (int64_t)t * p.n);
                am.x += a.x, am.y += a.y, am.z += a.z, am.w += a.w;
            }
        }
        const float4 rho4 = __ldg(reinterpret_cast<const float4*>(p.rho + i0));
        const float rho[4] = {rho4.x, rho4.y, rho4.z, rho4.w};
        float arv[4] = {ar.x, ar.y, ar.z, ar.w}, amv[4] = {am.x, am.y, am.z, am.w};
        if (KL) {
            const float4 mu4 = __ldg(reinterpret_cast<const float4*>(p.mu + i0));
            const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w};
            float sigma[4], pmu[4] = {0.f, 0.f, 0.f, 0.f}, inv_pvar[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) sigma[j] = bf_softplus(rho[j]);
            if (PRIOR == BF_PRIOR_GAUSSIAN) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p.prior_mu + i0));
                pmu[0] = v.x, pmu[1] = v.y, pmu[2] = v.z, pmu[3] = v.w;
                if (p.prior_rho != nullptr) {
                    const float4 pr = __ldg(reinterpret_cast<const float4*>(p.prior_rho + i0));
                    const float prho[4] = {pr.x, pr.y, pr.z, pr.w};
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
            float glq_sum = 0.0f;
            for (int s = 0; s < p.S; ++s) {
                float4 ev;
                if (p.eps_in != nullptr)
                    ev = __ldg(reinterpret_cast<const float4*>(p.eps_in + (int64_t)s * p.n + i0));
                else
                    ev = bf_eps_quad((uint32_t)q, (uint32_t)s, p.tensor_id, step, p.k0, p.k1);
                const float e[4] = {ev.x, ev.y, ev.z, ev.w};
                const float glq = p.g_logq ? __ldg(p.g_logq + s) : 0.0f;
                const float glp = p.g_logp ? __ldg(p.g_logp + s) : 0.0f;
                glq_sum += glq;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float w = __fadd_rn(mu[j], __fmul_rn(e[j], sigma[j]));
                    float dp = 0.0f;
                    if (PRIOR == BF_PRIOR_MIXTURE) dp = bf_mixture_dlogp(w, p.mix);
                    if (PRIOR == BF_PRIOR_GAUSSIAN) dp = -(w - pmu[j]) * inv_pvar[j];
                    const float gj = glp * dp;  // chain through w (both the mu and the sigma*eps path)
                    amv[j] += gj;
                    arv[j] += gj * e[j];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) arv[j] -= glq_sum / sigma[j];  // d log q / d sigma = -1/sigma, per sample
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) arv[j] *= bf_softplus_grad(rho[j]);
        float4 o = make_float4(arv[0], arv[1], arv[2], arv[3]);
        if (p.accumulate) {
            const float4 a = *reinterpret_cast<const float4*>(p.grad_rho + i0);
            o.x += a.x, o.y += a.y, o.z += a.z, o.w += a.w;
        }
        *reinterpret_cast<float4*>(p.grad_rho + i0) = o;
        if (p.grad_mu != nullptr) {
            float4 m = make_float4(amv[0], amv[1], amv[2], amv[3]);
            if (p.accumulate) {
                const float4 a = *reinterpret_cast<const float4*>(p.grad_mu + i0);
                m.x += a.x, m.y += a.y, m.z += a.z, m.w += a.w;
            }
            *reinterpret_cast<float4*>(p.grad_mu + i0) = m;
        }
    }
}

}  // namespace wg

int64_t bf_wgrad_fused_workspace_bytes_impl(int64_t S, int64_t M, int64_t N, int64_t K, int with_mu) {
    return S * wg::make_plan(S, M, N, K).splits * N * K * (int64_t)sizeof(float) * (with_mu ? 2 : 1);
}

int bf_linear_wgrad_fused_bf16(const void* gy, const void* x, int64_t S, int64_t M, int64_t N, int64_t K,
                               const float* mu, const float* rho, int32_t prior_kind, const float* prior_mu,
                               const float* prior_rho, float pi, float sigma1, float sigma2, const float* g_logq,
                               const float* g_logp, uint64_t seed, uint32_t step, uint32_t tensor_id,
                               const float* eps_in, float* grad_mu, float* grad_rho, int32_t accumulate,
                               void* workspace, cudaStream_t st) {
    using namespace wg;
    BF_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "bf16 path needs K % 8 == 0 and N % 8 == 0");
    BF_CHECK_ARG(workspace != nullptr, "workspace missing");
    BF_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 127u) == 0, "workspace must be 128 B aligned");
    const bool eps = eps_in != nullptr, with_mu = grad_mu != nullptr;
    Params p{};
    p.S = S, p.I = N, p.J = K, p.R = M;
    const Plan plan = make_plan(S, M, N, K);
    p.i_tiles = plan.i_tiles, p.j_tiles = plan.j_tiles, p.k_steps = plan.k_steps, p.splits = plan.splits;
    p.eps_in = eps_in;
    p.k0 = (uint32_t)(seed & 0xffffffffu), p.k1 = (uint32_t)(seed >> 32), p.step = step, p.tensor_id = tensor_id;
    p.step_ptr = bf_step_counter();
    const int64_t T = S * p.splits;
    float* const part_rho = reinterpret_cast<float*>(workspace);
    float* const part_mu = with_mu ? part_rho + T * N * K : nullptr;

    CUtensorMap ma, mb, mr, mm;
    int rc;
    if ((rc = tc::encode_map(&ma, gy, S, M, N, tc::BLOCK_K))) return rc;
    if ((rc = tc::encode_map(&mb, x, S, M, K, tc::BLOCK_K))) return rc;
    if ((rc = tc::encode_map(&mr, part_rho, T, N, K, BM, true))) return rc;
    if ((rc = tc::encode_map(&mm, with_mu ? part_mu : part_rho, T, N, K, BM, true))) return rc;
#define BF_WG_LAUNCH(PAIR)                                                                                        \
    (with_mu ? (eps ? launch<true, true, PAIR>(ma, mb, mr, mm, p, st) : launch<false, true, PAIR>(ma, mb, mr, mm, p, st)) \
             : (eps ? launch<true, false, PAIR>(ma, mb, mr, mm, p, st) : launch<false, false, PAIR>(ma, mb, mr, mm, p, st)))
    rc = plan.pair ? BF_WG_LAUNCH(true) : BF_WG_LAUNCH(false);
#undef BF_WG_LAUNCH
    if (rc) return rc;

    ReduceParams r{};
    r.part_rho = part_rho, r.part_mu = part_mu, r.T = (int)T;
    r.mu = mu, r.rho = rho, r.prior_mu = prior_mu, r.prior_rho = prior_rho, r.g_logq = g_logq, r.g_logp = g_logp;
    r.eps_in = eps_in, r.grad_mu = grad_mu, r.grad_rho = grad_rho, r.n = N * K, r.S = (int)S, r.accumulate = accumulate;
    r.k0 = p.k0, r.k1 = p.k1, r.step = step, r.tensor_id = tensor_id, r.step_ptr = p.step_ptr;
    r.prior_const_ipv = 1.0f / (sigma1 * sigma1);
    r.mix = bf_make_mixture(pi, sigma1, sigma2);
    const int64_t nquad = r.n >> 2;
    const int64_t want = (nquad + kRedThreads - 1) / kRedThreads, cap = (int64_t)bf_num_sms() * 8;
    const int grid = (int)(want < cap ? want : cap);
    const bool kl = g_logq != nullptr || g_logp != nullptr;
    if (!kl)
        wgrad_reduce_kernel<BF_PRIOR_NONE, false><<<grid, kRedThreads, 0, st>>>(r);
    else if (prior_kind == BF_PRIOR_MIXTURE)
        wgrad_reduce_kernel<BF_PRIOR_MIXTURE, true><<<grid, kRedThreads, 0, st>>>(r);
    else if (prior_kind == BF_PRIOR_GAUSSIAN)
        wgrad_reduce_kernel<BF_PRIOR_GAUSSIAN, true><<<grid, kRedThreads, 0, st>>>(r);
    else
        wgrad_reduce_kernel<BF_PRIOR_NONE, true><<<grid, kRedThreads, 0, st>>>(r);
    BF_LAUNCH_OK();
    return 0;
}
