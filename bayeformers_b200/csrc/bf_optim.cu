// Fused global-norm gradient clipping + AdamW over all trainable tensors.
//
// Replaces the step either side of the variational path in the reference's training loop
// (examples/bert_glue.py:240-241):
//     torch.nn.utils.clip_grad_norm_(model.parameters(), MAX_GRAD_NORM)     # norm pass + scale pass
//     optimizer.step()                                                      # AdamW: reads p,g,m,v, writes p,m,v
// with two launches and no scale pass: (1) per-chunk sums of squares, (2) every block re-derives the global norm
// from the chunk partials in a fixed order (deterministic, identical in all blocks), forms the clip coefficient and
// applies AdamW to its chunks with the clipped gradient.  HBM-bound: 28 B per fp32 element (+4 B for the norm pass).
// Arithmetic follows torch.optim.AdamW (decoupled weight decay, bias correction), whose CPU implementation is the
// oracle in the tests.
#include "bf_common.cuh"

namespace {

constexpr int kOptThreads = 256;
constexpr int kOptChunk = 16384;  // elements per work item (64 per thread)

template <typename T>
__device__ __forceinline__ float4 opt_ld4(const T* p);
template <>
__device__ __forceinline__ float4 opt_ld4<float>(const float* p) {
    return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 opt_ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
}
template <typename T>
__device__ __forceinline__ void opt_st4(T* p, float4 v);
template <>
__device__ __forceinline__ void opt_st4<float>(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void opt_st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&lo);
    u.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = bf_warp_sum(v);
    __syncthreads();  // red may still be read from a previous call
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < kOptThreads / 32; ++w) t += red[w];
    return t;
}

template <typename T>
__device__ __forceinline__ float chunk_sumsq(const T* g, int64_t lo, int64_t hi, bool vec) {
    float acc = 0.0f;
    if (vec) {
        for (int64_t i = lo + 4 * (int64_t)threadIdx.x; i < hi; i += 4 * kOptThreads) {
            const float4 v = opt_ld4<T>(g + i);
            acc = fmaf(v.x, v.x, acc), acc = fmaf(v.y, v.y, acc), acc = fmaf(v.z, v.z, acc), acc = fmaf(v.w, v.w, acc);
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += kOptThreads) {
            const float v = bf_ld_as_float(g + i);
            acc = fmaf(v, v, acc);
        }
    }
    return acc;
}

__global__ void __launch_bounds__(kOptThreads) grad_sumsq_kernel(const bf_opt_desc* __restrict__ descs,
                                                                 const int2* __restrict__ chunks, int n_chunks,
                                                                 float* __restrict__ partials,
                                                                 float* __restrict__ steps) {
    __shared__ float red[kOptThreads / 32];
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int2 ch = __ldg(chunks + c);
        const bf_opt_desc d = descs[ch.x];
        const int64_t lo = (int64_t)ch.y * kOptChunk;
        const int64_t hi = lo + kOptChunk < d.n ? lo + kOptChunk : d.n;
        float acc = 0.0f;
        if (d.grad != nullptr) {
            if (d.dtype == BF_BF16) acc = chunk_sumsq(reinterpret_cast<const __nv_bfloat16*>(d.grad), lo, hi, d.vec != 0);
            else acc = chunk_sumsq(reinterpret_cast<const float*>(d.grad), lo, hi, d.vec != 0);
        }
        const float t = block_sum(acc, red);
        if (threadIdx.x == 0) {
            partials[c] = t;
            // per-tensor step count, like torch (a tensor without a gradient does not advance): bumped once, by
            // the block that owns the tensor's first chunk; the update kernel reads it afterwards
            if (ch.y == 0 && d.grad != nullptr) steps[ch.x] += 1.0f;
        }
    }
}

struct AdamScalars {
    float lr, beta1, beta2, eps, weight_decay, max_norm;
};

template <typename T>
__device__ __forceinline__ void adamw_chunk(const bf_opt_desc& d, int64_t lo, int64_t hi, float clip, float decay,
                                            float step_size, float inv_sqrt_bc2, const AdamScalars& a) {
    T* const p = reinterpret_cast<T*>(d.param);
    const T* const g = reinterpret_cast<const T*>(d.grad);
    float* const master = d.master;  // fp32 master copy of a bf16 parameter (nullable): the update runs on it
    float* const sig = d.sigma_out;  // sigma cache of a rho tensor (nullable): softplus of the updated value
    auto upd = [&](float pv, float gv, float& m, float& v) {
        gv *= clip;
        pv *= decay;                               // p.mul_(1 - lr*wd)
        m = fmaf(gv - m, 1.0f - a.beta1, m);       // exp_avg.lerp_(g, 1 - beta1)
        v = fmaf(v, a.beta2, (1.0f - a.beta2) * gv * gv);
        const float denom = sqrtf(v) * inv_sqrt_bc2 + a.eps;
        return pv - step_size * (m / denom);
    };
    if (d.vec) {
        for (int64_t i = lo + 4 * (int64_t)threadIdx.x; i < hi; i += 4 * kOptThreads) {
            float4 pv = master ? *reinterpret_cast<const float4*>(master + i) : opt_ld4<T>(p + i);
            const float4 gv = opt_ld4<T>(g + i);
            float4 m = *reinterpret_cast<const float4*>(d.exp_avg + i);
            float4 v = *reinterpret_cast<const float4*>(d.exp_avg_sq + i);
            pv.x = upd(pv.x, gv.x, m.x, v.x), pv.y = upd(pv.y, gv.y, m.y, v.y);
            pv.z = upd(pv.z, gv.z, m.z, v.z), pv.w = upd(pv.w, gv.w, m.w, v.w);
            opt_st4<T>(p + i, pv);
            if (master) *reinterpret_cast<float4*>(master + i) = pv;
            if (sig)
                *reinterpret_cast<float4*>(sig + i) =
                    make_float4(bf_softplus(pv.x), bf_softplus(pv.y), bf_softplus(pv.z), bf_softplus(pv.w));
            *reinterpret_cast<float4*>(d.exp_avg + i) = m;
            *reinterpret_cast<float4*>(d.exp_avg_sq + i) = v;
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += kOptThreads) {
            float m = d.exp_avg[i], v = d.exp_avg_sq[i];
            const float pn = upd(master ? master[i] : bf_ld_as_float(p + i), bf_ld_as_float(g + i), m, v);
            if (sizeof(T) == 2) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(pn);
            else reinterpret_cast<float*>(p)[i] = pn;
            if (master) master[i] = pn;
            if (sig) sig[i] = bf_softplus(pn);
            d.exp_avg[i] = m, d.exp_avg_sq[i] = v;
        }
    }
}

__global__ void __launch_bounds__(kOptThreads) clip_adamw_kernel(const bf_opt_desc* __restrict__ descs,
                                                                 const int2* __restrict__ chunks, int n_chunks,
                                                                 const float* __restrict__ partials,
                                                                 const float* __restrict__ steps, AdamScalars a,
                                                                 float* __restrict__ grad_norm_out) {
    __shared__ float red[kOptThreads / 32];
    // global gradient norm: same fixed-order sum in every block
    float acc = 0.0f;
    for (int c = threadIdx.x; c < n_chunks; c += kOptThreads) acc += partials[c];
    const float total = sqrtf(block_sum(acc, red));
    if (blockIdx.x == 0 && threadIdx.x == 0 && grad_norm_out != nullptr) *grad_norm_out = total;
    float clip = 1.0f;
    if (a.max_norm > 0.0f) {
        clip = a.max_norm / (total + 1e-6f);  // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max=1)
        if (clip > 1.0f) clip = 1.0f;
    }
    const float decay = 1.0f - a.lr * a.weight_decay;
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int2 ch = __ldg(chunks + c);
        const bf_opt_desc d = descs[ch.x];
        if (d.grad == nullptr) continue;
        const float t = steps[ch.x];
        const float bc1 = 1.0f - powf(a.beta1, t), bc2 = 1.0f - powf(a.beta2, t);
        const float step_size = a.lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
        const int64_t lo = (int64_t)ch.y * kOptChunk;
        const int64_t hi = lo + kOptChunk < d.n ? lo + kOptChunk : d.n;
        if (d.dtype == BF_BF16) adamw_chunk<__nv_bfloat16>(d, lo, hi, clip, decay, step_size, inv_sqrt_bc2, a);
        else adamw_chunk<float>(d, lo, hi, clip, decay, step_size, inv_sqrt_bc2, a);
    }
}

}  // namespace

extern "C" int32_t bf_optim_chunk_elems(void) { return kOptChunk; }

extern "C" int64_t bf_clip_adamw_workspace_bytes(int64_t n_chunks) {
    return (n_chunks < 1 ? 1 : n_chunks) * (int64_t)sizeof(float);
}

extern "C" int bf_clip_adamw_step(const bf_opt_desc* descs, const int32_t* chunks, int32_t n_chunks, float lr,
                                  float beta1, float beta2, float eps, float weight_decay, float max_norm,
                                  float* steps, float* grad_norm_out, void* workspace, void* stream) {
    BF_CHECK_ARG(descs && chunks && steps && workspace, "null pointer");
    BF_CHECK_ARG(n_chunks >= 1, "no work");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    float* partials = reinterpret_cast<float*>(workspace);
    const int64_t cap = (int64_t)bf_num_sms() * 8;
    const int grid = (int)(n_chunks < cap ? n_chunks : cap);
    grad_sumsq_kernel<<<grid, kOptThreads, 0, st>>>(descs, reinterpret_cast<const int2*>(chunks), n_chunks, partials,
                                                    steps);
    BF_LAUNCH_OK();
    AdamScalars a{lr, beta1, beta2, eps, weight_decay, max_norm};
    clip_adamw_kernel<<<grid, kOptThreads, 0, st>>>(descs, reinterpret_cast<const int2*>(chunks), n_chunks, partials,
                                                    steps, a, grad_norm_out);
    BF_LAUNCH_OK();
    return 0;
}
