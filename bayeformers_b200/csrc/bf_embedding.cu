// S-sample Bayesian embedding lookup and its row-sparse backward.
//
// The reference snapshot has no bnn.Embedding; the north star specifies it by analogy with bnn.Linear: the whole
// table is a Gaussian (bayeformers/nn/parameters/gaussian.py:22-116), every forward samples it, reduces log q / log p
// over ALL of it, then looks rows up (F.embedding).  Materialising S sampled tables costs S*V*H*4 bytes per forward
// (BERT-large, S = 16: 2 GB) although only the looked-up rows are ever used.  eps is a pure function of
// (seed, step, tensor, sample, element), so row r of sample s needs no neighbours:
//
//   forward : out[t][:] = mu[id_t][:] + softplus(rho[id_t][:]) * eps_s(id_t*H + :),   s = t / tok_per_sample
//             (the log-prob sums over the whole table come from bf_sample_kl_fwd with w_out == NULL)
//   backward: grad_mu[r]  += sum_{t: id_t = r} g[t]
//             grad_rho[r] += sigmoid(rho[r]) * sum_{t: id_t = r} g[t] * eps_{s(t)}(r*H + :)
//
// The backward is deterministic without float atomics: the caller hands the tokens sorted by row id (index
// bookkeeping: one stable sort), blocks walk fixed chunks of the sorted list, a run of equal ids that lies inside one
// chunk is reduced and written by that block alone, runs that cross chunk borders leave per-chunk partials which one
// block per run then adds in chunk order.
#include "bf_common.cuh"

namespace {

constexpr int kEmbThreads = 256;
constexpr int kEmbChunk = 256;  // sorted tokens per block of the backward

struct EmbFwdParams {
    const int64_t* ids;
    const float* mu;
    const float* rho;
    const float* eps_in;  // [S][V*H] or null
    void* out;
    int64_t n_tok, tok_per_sample, V, H;
    uint32_t k0, k1, step, tensor_id;
    const uint32_t* step_ptr;
};

template <typename OT>
__device__ __forceinline__ void emb_store4(OT* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void emb_store4<float>(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void emb_store4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 v;
    v.x = *reinterpret_cast<const uint32_t*>(&lo);
    v.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = v;
}

// eps of the 4 elements [e0, e0+4) of the table (e0 % 4 == 0), sample s
__device__ __forceinline__ float4 emb_eps4(const float* eps_in, int64_t table_elems, int64_t e0, uint32_t s,
                                           uint32_t tensor_id, uint32_t step, uint32_t k0, uint32_t k1) {
    if (eps_in != nullptr) return __ldg(reinterpret_cast<const float4*>(eps_in + (int64_t)s * table_elems + e0));
    return bf_eps_quad((uint32_t)(e0 >> 2), s, tensor_id, step, k0, k1);
}

// H % 4 == 0: one thread per (token, column quad)
template <typename OT>
__global__ void __launch_bounds__(kEmbThreads) embedding_fwd_kernel(const EmbFwdParams p) {
    const uint32_t step = p.step + (p.step_ptr ? __ldg(p.step_ptr) : 0u);
    const int64_t hq = p.H >> 2;
    const int64_t total = p.n_tok * hq;
    OT* const out = reinterpret_cast<OT*>(p.out);
    for (int64_t i = (int64_t)blockIdx.x * kEmbThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kEmbThreads) {
        const int64_t t = i / hq, c = (i - t * hq) << 2;
        const int64_t id = __ldg(p.ids + t);
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        if ((uint64_t)id < (uint64_t)p.V) {
            const int64_t e0 = id * p.H + c;
            const float4 m = __ldg(reinterpret_cast<const float4*>(p.mu + e0));
            const float4 r = __ldg(reinterpret_cast<const float4*>(p.rho + e0));
            const float4 e = emb_eps4(p.eps_in, p.V * p.H, e0, (uint32_t)(t / p.tok_per_sample), p.tensor_id, step, p.k0,
                                      p.k1);
            // w = mu + eps*sigma with the reference's two roundings (gaussian.py:101), as in bf_sample_kl_fwd
            w[0] = __fadd_rn(m.x, __fmul_rn(e.x, bf_softplus(r.x)));
            w[1] = __fadd_rn(m.y, __fmul_rn(e.y, bf_softplus(r.y)));
            w[2] = __fadd_rn(m.z, __fmul_rn(e.z, bf_softplus(r.z)));
            w[3] = __fadd_rn(m.w, __fmul_rn(e.w, bf_softplus(r.w)));
        }
        emb_store4<OT>(out + t * p.H + c, w[0], w[1], w[2], w[3]);
    }
}

struct EmbBwdParams {
    const void* g;
    const int64_t* sorted_ids;
    const int64_t* perm;
    const float* rho;
    const float* eps_in;
    float* grad_mu;   // nullable
    float* grad_rho;
    float* part_mu;   // [n_chunks][2][H] (only when grad_mu)
    float* part_rho;  // [n_chunks][2][H]
    int64_t n_tok, tok_per_sample, V, H, padding_idx;
    int n_chunks;
    uint32_t k0, k1, step, tensor_id;
    const uint32_t* step_ptr;
};

__device__ __forceinline__ float4 emb_ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 emb_ld4(const __nv_bfloat16* p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
}

// add a finished row sum to the dense gradients (the only writer of this row: deterministic)
__device__ __forceinline__ void emb_commit(const EmbBwdParams& p, int64_t row, int64_t c, const float (&am)[4],
                                           const float (&ar)[4]) {
    if (row == p.padding_idx || (uint64_t)row >= (uint64_t)p.V) return;
    const int64_t e0 = row * p.H + c;
    const float4 r = emb_ld4(p.rho + e0);
    float4 gr = *reinterpret_cast<const float4*>(p.grad_rho + e0);
    gr.x += ar[0] * bf_softplus_grad(r.x), gr.y += ar[1] * bf_softplus_grad(r.y);
    gr.z += ar[2] * bf_softplus_grad(r.z), gr.w += ar[3] * bf_softplus_grad(r.w);
    *reinterpret_cast<float4*>(p.grad_rho + e0) = gr;
    if (p.grad_mu != nullptr) {
        float4 gm = *reinterpret_cast<const float4*>(p.grad_mu + e0);
        gm.x += am[0], gm.y += am[1], gm.z += am[2], gm.w += am[3];
        *reinterpret_cast<float4*>(p.grad_mu + e0) = gm;
    }
}

// stage 1: block b reduces the sorted tokens [b*kEmbChunk, (b+1)*kEmbChunk); thread owns column quads
template <typename GT>
__global__ void __launch_bounds__(kEmbThreads) embedding_bwd_kernel(const EmbBwdParams p) {
    __shared__ int64_t s_ids[kEmbChunk + 2];  // [0] = id before the chunk, [1..n] the chunk, [n+1] = id after it
    __shared__ int64_t s_tok[kEmbChunk];
    const uint32_t step = p.step + (p.step_ptr ? __ldg(p.step_ptr) : 0u);
    const GT* const g = reinterpret_cast<const GT*>(p.g);
    const int64_t c0 = (int64_t)blockIdx.x * kEmbChunk;
    const int64_t c1 = c0 + kEmbChunk < p.n_tok ? c0 + kEmbChunk : p.n_tok;
    const int n = (int)(c1 - c0);
    for (int j = threadIdx.x; j < n + 2; j += kEmbThreads) {
        const int64_t at = c0 - 1 + j;
        s_ids[j] = (at >= 0 && at < p.n_tok) ? __ldg(p.sorted_ids + at) : -1;  // ids are >= 0: -1 never matches
        if (j < n) s_tok[j] = __ldg(p.perm + c0 + j);
    }
    __syncthreads();
    const int64_t before = s_ids[0], after = s_ids[n + 1];
    const int64_t hq = p.H >> 2;
    const int64_t table = p.V * p.H;
    for (int64_t q = threadIdx.x; q < hq; q += kEmbThreads) {
        const int64_t c = q << 2;
        float am[4] = {0.f, 0.f, 0.f, 0.f}, ar[4] = {0.f, 0.f, 0.f, 0.f};
        int run_start = 0;
        int64_t row = s_ids[1];
        auto flush = [&](int j_end) {  // the run [run_start, j_end) of `row` is complete
            const bool left_open = run_start == 0 && before == row;
            const bool right_open = j_end == n && after == row;
            if (!left_open && !right_open) {
                emb_commit(p, row, c, am, ar);
            } else {
                const int slot = left_open ? 0 : 1;  // a run open on both sides is the chunk's only run: slot 0
                const int64_t o = ((int64_t)blockIdx.x * 2 + slot) * p.H + c;
                *reinterpret_cast<float4*>(p.part_rho + o) = make_float4(ar[0], ar[1], ar[2], ar[3]);
                if (p.grad_mu != nullptr)
                    *reinterpret_cast<float4*>(p.part_mu + o) = make_float4(am[0], am[1], am[2], am[3]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) am[k] = ar[k] = 0.0f;
        };
        for (int j0 = 0; j0 < n; j0 += 4) {
            float4 gv[4];  // 4 gradient rows in flight
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (j0 + u < n) gv[u] = emb_ld4(g + s_tok[j0 + u] * p.H + c);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                if (j >= n) break;
                const int64_t rj = s_ids[1 + j];
                if (rj != row) {
                    flush(j);
                    row = rj, run_start = j;
                }
                float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((uint64_t)row < (uint64_t)p.V)
                    e = emb_eps4(p.eps_in, table, row * p.H + c, (uint32_t)(s_tok[j] / p.tok_per_sample), p.tensor_id,
                                 step, p.k0, p.k1);
                am[0] += gv[u].x, am[1] += gv[u].y, am[2] += gv[u].z, am[3] += gv[u].w;
                ar[0] = fmaf(gv[u].x, e.x, ar[0]), ar[1] = fmaf(gv[u].y, e.y, ar[1]);
                ar[2] = fmaf(gv[u].z, e.z, ar[2]), ar[3] = fmaf(gv[u].w, e.w, ar[3]);
            }
        }
        flush(n);
    }
}

// stage 2: block b owns the run that STARTS in chunk b and continues past its end; it adds the partials of that run
// in chunk order (slot 1 of chunk b, then slot 0 of chunks b+1, b+2, ... while they continue the row)
__global__ void __launch_bounds__(kEmbThreads) embedding_bwd_fixup_kernel(const EmbBwdParams p) {
    const int b = blockIdx.x;
    const int64_t c0 = (int64_t)b * kEmbChunk;
    const int64_t c1 = c0 + kEmbChunk < p.n_tok ? c0 + kEmbChunk : p.n_tok;
    if (c1 >= p.n_tok) return;
    const int64_t row = __ldg(p.sorted_ids + c1 - 1);
    if (__ldg(p.sorted_ids + c1) != row) return;  // last run closed at the border
    // head of the run: it did not already come in from the left over the whole chunk
    const bool whole = __ldg(p.sorted_ids + c0) == row;
    if (whole && c0 > 0 && __ldg(p.sorted_ids + c0 - 1) == row) return;
    // chunks b+1 .. e-1 continue the row with their slot-0 run
    int e = b + 1;
    while (e < p.n_chunks) {
        const int64_t d0 = (int64_t)e * kEmbChunk;
        const int64_t d1 = d0 + kEmbChunk < p.n_tok ? d0 + kEmbChunk : p.n_tok;
        ++e;
        // does the row run through the whole of this chunk and beyond?
        if (!(__ldg(p.sorted_ids + d1 - 1) == row && d1 < p.n_tok && __ldg(p.sorted_ids + d1) == row)) break;
    }
    const int64_t hq = p.H >> 2;
    for (int64_t q = threadIdx.x; q < hq; q += kEmbThreads) {
        const int64_t c = q << 2;
        float am[4], ar[4];
        {
            const int64_t o = ((int64_t)b * 2 + 1) * p.H + c;
            const float4 r = *reinterpret_cast<const float4*>(p.part_rho + o);
            ar[0] = r.x, ar[1] = r.y, ar[2] = r.z, ar[3] = r.w;
            am[0] = am[1] = am[2] = am[3] = 0.0f;
            if (p.grad_mu != nullptr) {
                const float4 m = *reinterpret_cast<const float4*>(p.part_mu + o);
                am[0] = m.x, am[1] = m.y, am[2] = m.z, am[3] = m.w;
            }
        }
#pragma unroll 4
        for (int k = b + 1; k < e; ++k) {
            const int64_t o = ((int64_t)k * 2) * p.H + c;
            const float4 r = *reinterpret_cast<const float4*>(p.part_rho + o);
            ar[0] += r.x, ar[1] += r.y, ar[2] += r.z, ar[3] += r.w;
            if (p.grad_mu != nullptr) {
                const float4 m = *reinterpret_cast<const float4*>(p.part_mu + o);
                am[0] += m.x, am[1] += m.y, am[2] += m.z, am[3] += m.w;
            }
        }
        emb_commit(p, row, c, am, ar);
    }
}

inline int emb_chunks(int64_t n_tok) { return (int)((n_tok + kEmbChunk - 1) / kEmbChunk); }

}  // namespace

extern "C" int bf_embedding_supported(int64_t H) { return (H >= 4 && H % 4 == 0) ? 1 : 0; }

extern "C" int bf_embedding_fwd(const int64_t* ids, int64_t n_tok, int64_t tok_per_sample, const float* mu,
                                const float* rho, int64_t V, int64_t H, uint64_t seed, uint32_t step,
                                uint32_t tensor_id, const float* eps_in, void* out, int32_t out_dtype, void* stream) {
    BF_CHECK_ARG(n_tok >= 0 && tok_per_sample >= 1 && V >= 1, "bad sizes");
    BF_CHECK_ARG(bf_embedding_supported(H), "embedding width must be a multiple of 4");
    BF_CHECK_ARG(out_dtype == BF_F32 || out_dtype == BF_BF16, "bad out_dtype");
    if (n_tok == 0) return 0;
    BF_CHECK_ARG(ids && mu && rho && out, "null pointer");
    BF_CHECK_ARG(((reinterpret_cast<uintptr_t>(mu) | reinterpret_cast<uintptr_t>(rho) | reinterpret_cast<uintptr_t>(out) |
                   reinterpret_cast<uintptr_t>(eps_in)) & 15u) == 0, "mu, rho, eps_in, out must be 16 B aligned");
    EmbFwdParams p{};
    p.ids = ids, p.mu = mu, p.rho = rho, p.eps_in = eps_in, p.out = out;
    p.n_tok = n_tok, p.tok_per_sample = tok_per_sample, p.V = V, p.H = H;
    p.k0 = (uint32_t)(seed & 0xffffffffu), p.k1 = (uint32_t)(seed >> 32), p.step = step, p.tensor_id = tensor_id;
    p.step_ptr = bf_step_counter();
    const int64_t total = n_tok * (H >> 2);
    const int64_t want = (total + kEmbThreads - 1) / kEmbThreads, cap = (int64_t)bf_num_sms() * 16;
    const int grid = (int)(want < cap ? want : cap);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (out_dtype == BF_BF16) embedding_fwd_kernel<__nv_bfloat16><<<grid, kEmbThreads, 0, st>>>(p);
    else embedding_fwd_kernel<float><<<grid, kEmbThreads, 0, st>>>(p);
    BF_LAUNCH_OK();
    return 0;
}

extern "C" int64_t bf_embedding_bwd_workspace_bytes(int64_t n_tok, int64_t H) {
    const int64_t chunks = n_tok < 1 ? 1 : emb_chunks(n_tok);
    return 2 * chunks * 2 * H * (int64_t)sizeof(float);  // (mu, rho) x [chunks][2][H]
}

extern "C" int bf_embedding_bwd(const void* g, int32_t g_dtype, const int64_t* sorted_ids, const int64_t* perm,
                                int64_t n_tok, int64_t tok_per_sample, const float* rho, int64_t V, int64_t H,
                                int64_t padding_idx, uint64_t seed, uint32_t step, uint32_t tensor_id,
                                const float* eps_in, float* grad_mu, float* grad_rho, void* workspace, void* stream) {
    BF_CHECK_ARG(n_tok >= 0 && tok_per_sample >= 1 && V >= 1, "bad sizes");
    BF_CHECK_ARG(bf_embedding_supported(H), "embedding width must be a multiple of 4");
    BF_CHECK_ARG(g_dtype == BF_F32 || g_dtype == BF_BF16, "bad g_dtype");
    if (n_tok == 0) return 0;
    BF_CHECK_ARG(g && sorted_ids && perm && rho && grad_rho && workspace, "null pointer");
    BF_CHECK_ARG(((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(rho) | reinterpret_cast<uintptr_t>(grad_rho) |
                   reinterpret_cast<uintptr_t>(grad_mu) | reinterpret_cast<uintptr_t>(eps_in) |
                   reinterpret_cast<uintptr_t>(workspace)) & 15u) == 0, "buffers must be 16 B aligned");
    EmbBwdParams p{};
    p.g = g, p.sorted_ids = sorted_ids, p.perm = perm, p.rho = rho, p.eps_in = eps_in;
    p.grad_mu = grad_mu, p.grad_rho = grad_rho;
    p.n_tok = n_tok, p.tok_per_sample = tok_per_sample, p.V = V, p.H = H, p.padding_idx = padding_idx;
    p.n_chunks = emb_chunks(n_tok);
    p.part_rho = reinterpret_cast<float*>(workspace);
    p.part_mu = p.part_rho + (int64_t)p.n_chunks * 2 * H;
    p.k0 = (uint32_t)(seed & 0xffffffffu), p.k1 = (uint32_t)(seed >> 32), p.step = step, p.tensor_id = tensor_id;
    p.step_ptr = bf_step_counter();
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (g_dtype == BF_BF16) embedding_bwd_kernel<__nv_bfloat16><<<p.n_chunks, kEmbThreads, 0, st>>>(p);
    else embedding_bwd_kernel<float><<<p.n_chunks, kEmbThreads, 0, st>>>(p);
    BF_LAUNCH_OK();
    embedding_bwd_fixup_kernel<<<p.n_chunks, kEmbThreads, 0, st>>>(p);
    BF_LAUNCH_OK();
    return 0;
}
