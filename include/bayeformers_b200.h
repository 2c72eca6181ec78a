/*
 * bayeformers_b200 -- C ABI of the B200-native variational-layer hot path.
 *
 * The reference (yliess86/BayeFormers) has no FFI of its own: its operator
 * surface is the Python module `bayeformers.nn`.  This
 * header is the boundary a maintainer of the reference would bind to replace
 * the arithmetic inside
 *
 *     bayeformers/nn/parameters/gaussian.py:90-116   Gaussian.sample / log_prob
 *     bayeformers/nn/parameters/gaussian.py:160-171  ScaledGaussianMixture.log_prob
 *     bayeformers/nn/layers/linear.py:83-104         Linear.forward (+ its autograd)
 *
 * Conventions (all functions):
 *   - plain C types only: raw DEVICE pointers, sizes as int64_t, the CUDA
 *     stream as an opaque `void*` (a cudaStream_t);
 *   - return 0 on success, otherwise a non-zero code (a cudaError_t, or
 *     BF_ERR_*); `bf_last_error()` gives the text for the calling thread;
 *   - never allocate device memory, never synchronise the stream, never
 *     throw; workspaces are caller-provided and sized by the *_workspace_bytes
 *     queries; the caller keeps every buffer alive until the stream reaches
 *     the kernel;
 *   - a workspace must be zero-filled ONCE when it is allocated (it holds a
 *     self-resetting completion counter) and must not be shared by launches
 *     that may run concurrently on different streams.
 *
 * Philox contract (what "eps" is when `eps_in == NULL`): see oracle/philox_oracle.py
 * and DESIGN.md.  counter = (flat_index >> 2, sample_id, tensor_id, step),
 * key = seed; philox4x32-10; two Box-Muller pairs per counter.
 */
#ifndef BAYEFORMERS_B200_H_
#define BAYEFORMERS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BF_ABI_VERSION 2

/* element types of activation / sampled-weight buffers */
#define BF_F32 0
#define BF_BF16 1

/* prior kinds */
#define BF_PRIOR_MIXTURE 0  /* ScaledGaussianMixture(pi, sigma1, sigma2): gaussian.py:119-171 */
#define BF_PRIOR_GAUSSIAN 1 /* Gaussian(prior_mu, prior_rho), the MOPED prior: linear.py:147-150 */
#define BF_PRIOR_NONE 2     /* no prior term (log p contribution 0) */

/* error codes beyond cudaError_t */
#define BF_ERR_BAD_ARG 10001
#define BF_ERR_UNSUPPORTED 10002
#define BF_ERR_DRIVER 10003

int bf_abi_version(void);
const char* bf_last_error(void);
/* 1 when the running device is sm_100 (tcgen05 path usable), else 0; <0 on error */
int bf_device_is_sm100(void);
/* Optional device-resident step counter, registered for the CURRENT device (cudaGetDevice at
 * the time of the call; one slot per device): when set, every kernel launched on that device
 * that draws eps uses  step + *device_counter  as the Philox step.  This is what makes a
 * whole training step capturable in a CUDA graph: the `step` arguments get baked into the
 * graph, the counter is bumped on the device between replays.  NULL switches it off. */
int bf_set_step_counter(const uint32_t* device_counter);

/* Tuning switches for A/B measurements.  The library never reads the environment; a caller that
 * wants a non-default kernel choice says so here.  Defaults are the production choices. */
#define BF_OPT_GEMM_2CTA 0         /* fwd/dgrad: 0 single-CTA kernels, 1 auto (default), 2 force CTA pairs */
#define BF_OPT_WGRAD_2CTA 1        /* fused wgrad: same values */
#define BF_OPT_RESLN_BWD_STAGED 2  /* 1 (default): shared-memory-staged resln backward, 0: register prefetch */
#define BF_OPT_SK_PREFETCH 3       /* multi-tensor sample+KL prefetch: 0 none, 1 L1 (default), 2 L2 */
#define BF_OPT_ATTN_TC 4          /* T == 128 attention: 1 (default) tcgen05 kernels, 0 the mma.sync kernels */
#define BF_OPT_GELU_POLY 5        /* fused GELU / GELU' epilogues: 1 (default) odd polynomials on FFMA2 (|err| <= 1e-4 /
                                    6e-4, far inside bf16 rounding), 0 the erf forms (A&S 7.1.26 / 7.1.28) */
#define BF_OPT_COUNT 6
int bf_set_option(int32_t option, int32_t value);
int bf_get_option(int32_t option);

/* ------------------------------------------------------------------------- *
 * eps stream, exposed for the statistical tests.
 * Replaces: eps = self.normal.sample(self.size)   (gaussian.py:100)
 * out[i] = eps(i | seed, step, tensor_id, sample_id), i in [0, n)
 * ------------------------------------------------------------------------- */
int bf_philox_normal(float* out, int64_t n, uint64_t seed, uint32_t step, uint32_t tensor_id,
                     uint32_t sample_id, void* stream);

/* ------------------------------------------------------------------------- *
 * Fused sample + log q + log p for one variational tensor and S MC samples.
 * Replaces, per sample: Gaussian.sample (gaussian.py:90-101), Gaussian.log_prob
 * of the posterior (gaussian.py:103-116) and the prior's log_prob
 * (gaussian.py:160-171 mixture, or gaussian.py:103-116 on the MOPED prior) as
 * called from Linear.forward (linear.py:97-102).
 *
 *   for s in [0,S):  eps_s = eps_in ? eps_in[s*n + i] : philox(seed, step, tensor_id, s)
 *                    w_s[i] = mu[i] + eps_s[i] * softplus(rho[i])      -> w_out[s*w_stride + i]
 *                    logq[s] (+)= sum_i  -log sqrt(2pi) - log sigma_i - (w_s[i]-mu[i])^2 / (2 sigma_i^2)
 *                    logp[s] (+)= sum_i  log p(w_s[i])
 *
 * mu, rho            [n] fp32
 * prior_mu, prior_rho [n] fp32 (BF_PRIOR_GAUSSIAN only; prior_mu may alias mu).
 *                    prior_rho == NULL means a CONSTANT prior sigma, passed in `sigma1`
 *                    (MOPED sets rho_p = 1 everywhere: linear.py:149 -> sigma_p = softplus(1));
 *                    saves 4 B/element of reads and one softplus+log per element
 * eps_in             NULL, or [S*n] fp32 injected eps (parity tests)
 * w_out              [S][w_stride] of w_dtype (BF_F32 / BF_BF16); may be NULL (log-probs only)
 * logq_out, logp_out [S] fp32; `accumulate` != 0 adds to the existing values
 *                    (weight then bias, linear.py:99-102)
 * workspace          bf_sample_kl_workspace_bytes(n, S) bytes, zero-filled once
 * ------------------------------------------------------------------------- */
int64_t bf_sample_kl_workspace_bytes(int64_t n, int32_t S);
int bf_sample_kl_fwd(const float* mu, const float* rho, int32_t prior_kind, const float* prior_mu,
                     const float* prior_rho, float pi, float sigma1, float sigma2, int64_t n, int32_t S,
                     uint64_t seed, uint32_t step, uint32_t tensor_id, const float* eps_in, void* w_out,
                     int32_t w_dtype, int64_t w_stride, float* logq_out, float* logp_out, int32_t accumulate,
                     void* workspace, void* stream);

/* ------------------------------------------------------------------------- *
 * Multi-tensor form of bf_sample_kl_fwd: all variational tensors of a model in
 * one launch per sample chunk.  Same arithmetic,
 * same eps stream, same results as calling bf_sample_kl_fwd tensor by tensor
 * (log-prob sums agree to fp32 summation order).
 *
 * descs        DEVICE array of n_tensors descriptors (below)
 * chunks       DEVICE int32 pairs {tensor index, first quad}: the work list, each entry
 *              covering at most bf_sample_kl_multi_chunk_quads() quads (4 elements) of
 *              one tensor, entries of one output slot contiguous
 * slot_ranges  DEVICE int32 pairs {first chunk, last chunk + 1} per output slot; a slot
 *              is what one (logq, logp) pair sums over -- typically weight + bias of a layer
 * prior_mask   bit k set when some descriptor has prior_kind == k (the descriptors live on the device; one kernel
 *              instantiation per prior kind present is launched over the chunk list)
 * logq_out, logp_out  [n_slots][S] fp32 (overwritten)
 * step         added to every descriptor's own `step`
 * w_base       NULL, or a base address: descriptors' `w_out` are then byte OFFSETS from it
 *              (lets a static descriptor table serve a freshly allocated output arena per call)
 * workspace    bf_sample_kl_multi_workspace_bytes(n_chunks) bytes, contents irrelevant
 * ------------------------------------------------------------------------- */
typedef struct bf_tensor_desc {
    const float* mu;
    const float* rho;
    const float* prior_mu;  /* BF_PRIOR_GAUSSIAN only */
    const float* prior_rho; /* NULL: constant prior sigma in `sigma1` */
    void* w_out;            /* [S][w_stride] of w_dtype; NULL (without w_base) or (void*)-1: log-probs only */
    int64_t n;
    int64_t w_stride;
    uint32_t tensor_id;
    uint32_t step;
    int32_t prior_kind;
    int32_t w_dtype;
    float pi, sigma1, sigma2;
    int32_t vec; /* 1: all pointers 16 B aligned, n % 4 == 0, (w_stride * elem) % 16 == 0 */
    const float* sigma; /* NULL, or [n] fp32 = softplus(rho) kept up to date by the caller (bf_clip_adamw_step writes it
                           through bf_opt_desc.sigma_out, bf_softplus_fwd fills it): read INSTEAD of rho, which spares
                           the exp / log1p of the per-element prologue (vector path only; same bits as computing it) */
} bf_tensor_desc;

/* sigma[i] = softplus(rho[i]) with the library's own softplus (initial fill of a sigma cache) */
int bf_softplus_fwd(const float* rho, float* sigma, int64_t n, void* stream);
int32_t bf_sample_kl_multi_chunk_quads(void);
int64_t bf_sample_kl_multi_workspace_bytes(int64_t n_chunks);
int bf_sample_kl_fwd_multi(const bf_tensor_desc* descs, const int32_t* chunks, int32_t n_chunks,
                           const int32_t* slot_ranges, int32_t n_slots, int32_t prior_mask, int32_t S, uint64_t seed,
                           uint32_t step, float* logq_out, float* logp_out, void* workspace, void* w_base, void* stream);

/* ------------------------------------------------------------------------- *
 * Backward of the above (stand-alone form): eps is RECOMPUTED from the seed.
 * Replaces what autograd does for gaussian.py:101 (mu + eps*softplus(rho)) and,
 * when g_logq/g_logp are given, for the two log_prob reductions (the KL
 * gradient the reference drops through `.data =`, linear.py:99-102).
 *
 *   grad_mu[i]  (+)= sum_s grad_w[s][i] + g_logp[s] * dlogp/dw(w_s[i])
 *   grad_rho[i] (+)= sigmoid(rho[i]) * sum_s ( grad_w[s][i]*eps_s[i]
 *                        - g_logq[s]/sigma_i + g_logp[s]*dlogp/dw(w_s[i])*eps_s[i] )
 *
 * grad_w   [S][gw_stride] of gw_dtype, or NULL (KL terms only)
 * g_logq, g_logp  device [S] fp32 upstream gradients of the two scalars, or NULL
 *                 (both NULL == the reference's behaviour: no KL gradient)
 * grad_mu  [n] fp32 or NULL (frozen mu);  grad_rho [n] fp32
 * ------------------------------------------------------------------------- */
int bf_sample_kl_bwd(const void* grad_w, int32_t gw_dtype, int64_t gw_stride, const float* mu, const float* rho,
                     int32_t prior_kind, const float* prior_mu, const float* prior_rho, float pi, float sigma1,
                     float sigma2, const float* g_logq, const float* g_logp, int64_t n, int32_t S, uint64_t seed,
                     uint32_t step, uint32_t tensor_id, const float* eps_in, float* grad_mu, float* grad_rho,
                     int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------- *
 * S-sample Linear contractions.  Replaces F.linear(input, weight, bias)
 * (linear.py:104) and its autograd (mm / addmm backward), batched over the
 * S Monte-Carlo weight samples.
 *
 * dtype == BF_F32 : fp32 FFMA kernels (reference-precision "parity" mode)
 * dtype == BF_BF16: tcgen05 / TMEM / TMA kernels, bf16 operands, fp32 accumulate
 *
 *   fwd   : y[s]  = x[s] (M,K) . w[s]^T (K,N) + bias[s]      y of y_dtype
 *   dgrad : dx[s] = gy[s] (M,N) . w[s] (N,K)                 dx of dx_dtype
 *   wgrad : dw[s] = gy[s]^T (N,M) . x[s] (M,K)               dw fp32 [S,N,K]
 *
 * bias  [S,N] fp32 or NULL.  All matrices row-major and densely packed.
 * ------------------------------------------------------------------------- */
int bf_linear_fwd(const void* x, const void* w, const float* bias, void* y, int64_t S, int64_t M, int64_t N,
                  int64_t K, int32_t dtype, int32_t y_dtype, void* stream);
int bf_linear_dgrad(const void* gy, const void* w, void* dx, int64_t S, int64_t M, int64_t N, int64_t K,
                    int32_t dtype, int32_t dx_dtype, void* stream);
int bf_linear_wgrad(const void* gy, const void* x, float* dw, int64_t S, int64_t M, int64_t N, int64_t K,
                    int32_t dtype, void* stream);
/* dx[s] += gy[s] . w[s]: the result tiles are added to dx by TMA reduce-add (cp.reduce.async.bulk.tensor .add, in
 * the element type of dx) instead of stored -- the gradient of an input that also feeds other consumers (the residual
 * branch, the q / k / v projections of one attention block) is accumulated in place, which replaces autograd's
 * separate add passes.  bf16 tensor-core path only (dtype == BF_BF16). */
int bf_linear_dgrad_accumulate(const void* gy, const void* w, void* dx, int64_t S, int64_t M, int64_t N, int64_t K,
                               int32_t dtype, int32_t dx_dtype, void* stream);

/* Split-precision ("fp32x3") contractions: reference precision (linear.py:104 computes in fp32, TF32 off) on the
 * tensor cores.  fp32 operands are given as bf16 (hi, lo) pairs, hi = bf16(v), lo = bf16(v - hi)
 * (bf_split_bf16x2); each output tile accumulates A_hi B_hi + A_hi B_lo + A_lo B_hi in one fp32 TMEM accumulator
 * (tcgen05, three passes over the reduction); the dropped A_lo B_lo term is ~2^-16 of a product.  Results fp32,
 * norm-wise error ~4e-6 of the fp32 result.  Shapes as bf_linear_fwd / _dgrad / _wgrad; N % 8 == 0 and K % 8 == 0. */
int bf_split_bf16x2(const float* src, void* hi, void* lo, int64_t n, void* stream);
int bf_linear_fwd_x3(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, float* y,
                     int64_t S, int64_t M, int64_t N, int64_t K, void* stream);
int bf_linear_dgrad_x3(const void* gy_hi, const void* gy_lo, const void* w_hi, const void* w_lo, float* dx, int64_t S,
                       int64_t M, int64_t N, int64_t K, void* stream);
int bf_linear_wgrad_x3(const void* gy_hi, const void* gy_lo, const void* x_hi, const void* x_lo, float* dw, int64_t S,
                       int64_t M, int64_t N, int64_t K, void* stream);

/* Extension: forward with the bias + GELU (exact, erf form) epilogue fused, for layers used as
 * y = gelu(F.linear(x, w, b)) (linear.py:104 followed by the host model's activation), and the matching
 * backward elementwise pass.  bf16 tensor-core path only, bias required.
 *   bf_linear_fwd_gelu      z[s] = x[s] w[s]^T + bias[s]  (bf16, kept for backward),  y[s] = gelu(z[s])  (bf16)
 *   bf_gelu_bwd_bias_grad   gz = gy * gelu'(z) (bf16),  db[s][j] = sum_m gz[s][m][j] (fp32; replaces bf_bias_grad)
 * bf_linear_fwd_gelu_supported: 1 when the CTA-pair kernel takes the shape (enough 256 x 256 tiles), else the
 * caller composes bf_linear_fwd with a separate GELU.  workspace of bf_gelu_bwd_bias_grad:
 * bf_gelu_bwd_bias_grad_workspace_bytes(S, M, N) bytes, zero-filled once. */
int bf_linear_fwd_gelu_supported(int64_t S, int64_t M, int64_t N, int64_t K);
int bf_linear_fwd_gelu(const void* x, const void* w, const float* bias, void* z, void* y, int64_t S, int64_t M,
                       int64_t N, int64_t K, void* stream);
/* dgrad of the Linear that CONSUMES a = gelu(z), with the activation's derivative in its epilogue:
 *   bf_linear_dgrad_gelu   gz[s] = (gy[s] . w[s]) o gelu'(z[s])    gy [S,M,N], w [S,N,K], z / gz [S,M,K], all bf16
 * i.e. bf_linear_dgrad followed by the elementwise half of bf_gelu_bwd_bias_grad in one kernel: the [S*M, K] gradient
 * crosses HBM once instead of three times (the z tile is TMA-loaded next to the operands).  The bias gradient of the
 * layer that produced z is then bf_bias_grad(gz).  bf_linear_dgrad_gelu_supported: 1 when the CTA-pair kernel takes
 * the shape, else the caller composes bf_linear_dgrad with bf_gelu_bwd_bias_grad. */
int bf_linear_dgrad_gelu_supported(int64_t S, int64_t M, int64_t N, int64_t K);
int bf_linear_dgrad_gelu(const void* gy, const void* w, const void* z, void* gz, int64_t S, int64_t M, int64_t N,
                         int64_t K, void* stream);
/* the same with the column sums of gz -- the bias gradient of the layer that produced z, dbias [S, K] fp32 -- taken from
 * the staged result boxes in the epilogue (per-block partial rows in `workspace`, added in block order by a second small
 * pass: deterministic), instead of a separate pass over gz.  dbias == NULL: plain bf_linear_dgrad_gelu. */
int64_t bf_linear_dgrad_gelu_bias_workspace_bytes(int64_t S, int64_t M, int64_t K);
int bf_linear_dgrad_gelu_bias(const void* gy, const void* w, const void* z, void* gz, float* dbias, void* workspace,
                              int64_t S, int64_t M, int64_t N, int64_t K, void* stream);
int64_t bf_gelu_bwd_bias_grad_workspace_bytes(int64_t S, int64_t M, int64_t N);
int bf_gelu_bwd_bias_grad(const void* gy, const void* z, void* gz, float* db, int64_t S, int64_t M, int64_t N,
                          void* workspace, void* stream);

/* wgrad with the variational backward fused into its epilogue: the raw weight
 * gradients of the S samples never reach HBM (unless mu is trainable).  The
 * tensor-core epilogue regenerates the eps tile of each (sample, output tile)
 * and emits dW_s o eps_s; a fixed-order reduction pass then forms
 *   grad_mu  (+)= sum_s dw[s]                        (skipped when grad_mu == NULL)
 *   grad_rho (+)= sigmoid(rho) * sum_s dw[s]*eps_s   (+ the KL terms as in bf_sample_kl_bwd)
 * mu/rho/prior/eps arguments as in bf_sample_kl_bwd, n == N*K.
 * workspace: bf_linear_wgrad_fused_workspace_bytes(S, M, N, K, grad_mu != NULL) bytes,
 * 128 B aligned, contents irrelevant (per-(sample, reduction-slice) partial products;
 * written and consumed inside the call).  dtype must be BF_BF16.  Deterministic:
 * no float atomics, partials are summed in a fixed order. */
int64_t bf_linear_wgrad_fused_workspace_bytes(int64_t S, int64_t M, int64_t N, int64_t K, int32_t with_grad_mu);
int bf_linear_wgrad_fused(const void* gy, const void* x, int64_t S, int64_t M, int64_t N, int64_t K, int32_t dtype,
                          const float* mu, const float* rho, int32_t prior_kind, const float* prior_mu,
                          const float* prior_rho, float pi, float sigma1, float sigma2, const float* g_logq,
                          const float* g_logp, uint64_t seed, uint32_t step, uint32_t tensor_id,
                          const float* eps_in, float* grad_mu, float* grad_rho, int32_t accumulate,
                          void* workspace, void* stream);

/* column sums of gy over the M rows of each sample: db[s][j] = sum_m gy[s][m][j]
 * (the bias gradient F.linear's autograd produces).  db fp32 [S,N].  Deterministic
 * two-stage reduction; workspace of bf_bias_grad_workspace_bytes(S, M, N) bytes,
 * zero-filled once. */
int64_t bf_bias_grad_workspace_bytes(int64_t S, int64_t M, int64_t N);
int bf_bias_grad(const void* gy, int32_t gy_dtype, float* db, int64_t S, int64_t M, int64_t N, void* workspace,
                 void* stream);

/* ------------------------------------------------------------------------- *
 * S-sample LayerNorm (the Bayesian LayerNorm the north_star
 * names, absent from the reference snapshot; specified as the reference's
 * Gaussian.sample (gaussian.py:90-101) composed with F.layer_norm).
 *
 *   fwd : y[s][m][:] = (x[s][m][:] - mean) * rstd * gamma[s][:] + beta[s][:]
 *   bwd : dx, dgamma[s][:] = sum_m gy*xhat, dbeta[s][:] = sum_m gy
 *
 * x, y, gy, dx  [S*M, H] of `dtype` (BF_F32 / BF_BF16), rows [s*M, (s+1)*M) use sample s
 * gamma, beta   fp32, sample s at offset s*affine_stride (affine_stride == 0: one
 *               shared affine -- the frequentist LayerNorm); beta may be NULL
 * mean, rstd    [S*M] fp32, written by fwd, read by bwd
 * H             256, 512, 768 or 1024 (bf_layernorm_supported); statistics in fp32
 * workspace     bf_layernorm_bwd_workspace_bytes(S, M, H) bytes, zero-filled once;
 *               deterministic two-stage reduction of dgamma / dbeta
 * ------------------------------------------------------------------------- */
int bf_layernorm_supported(int64_t H);
int bf_layernorm_fwd(const void* x, int32_t dtype, const float* gamma, const float* beta, int64_t affine_stride,
                     int64_t S, int64_t M, int64_t H, float eps, void* y, float* mean, float* rstd, void* stream);
int64_t bf_layernorm_bwd_workspace_bytes(int64_t S, int64_t M, int64_t H);
int bf_layernorm_bwd(const void* gy, const void* x, int32_t dtype, const float* gamma, int64_t affine_stride,
                     const float* mean, const float* rstd, int64_t S, int64_t M, int64_t H, void* dx, float* dgamma,
                     float* dbeta, void* workspace, void* stream);

/* ------------------------------------------------------------------------- *
 * S-sample Embedding (the Bayesian Embedding the north_star names, absent from
 * the reference snapshot; specified as the reference's Gaussian.sample (gaussian.py:90-101) of the
 * whole table composed with F.embedding).  Only the looked-up rows are sampled -- eps is a pure
 * function of (seed, step, tensor_id, sample, element), so the S sampled tables are never
 * materialised; the log-prob sums over the WHOLE table come from bf_sample_kl_fwd / _multi with
 * w_out == NULL, using the same (seed, step, tensor_id).
 *
 *   fwd : out[t][:] = mu[id_t][:] + softplus(rho[id_t][:]) * eps_s(id_t*H + :),  s = t / tok_per_sample
 *   bwd : grad_mu[r]  += sum_{t: id_t = r} g[t]
 *         grad_rho[r] += sigmoid(rho[r]) * sum_{t: id_t = r} g[t] * eps_{s(t)}(r*H + :)
 *
 * ids          [n_tok] int64 (row ids; out-of-range ids give zero rows and no gradient)
 * mu, rho      [V, H] fp32, H % 4 == 0 (bf_embedding_supported)
 * eps_in       NULL, or [S][V*H] fp32 injected eps (parity tests)
 * out, g       [n_tok, H] of out_dtype / g_dtype (BF_F32 / BF_BF16)
 * sorted_ids, perm  the ids in ascending order and, for each sorted position, the original token
 *              index (a stable sort by the caller -- index bookkeeping, no arithmetic)
 * grad_mu, grad_rho  [V, H] fp32, ACCUMULATED into: the caller initialises them (zeros, or the
 *              KL terms from bf_sample_kl_bwd with grad_w == NULL); grad_mu may be NULL (frozen mu);
 *              rows equal to padding_idx receive nothing.  Deterministic: every row is reduced in
 *              sorted-token order by one thread block, no float atomics.
 * workspace    bf_embedding_bwd_workspace_bytes(n_tok, H) bytes, contents irrelevant
 * ------------------------------------------------------------------------- */
int bf_embedding_supported(int64_t H);
int bf_embedding_fwd(const int64_t* ids, int64_t n_tok, int64_t tok_per_sample, const float* mu, const float* rho,
                     int64_t V, int64_t H, uint64_t seed, uint32_t step, uint32_t tensor_id, const float* eps_in,
                     void* out, int32_t out_dtype, void* stream);
int64_t bf_embedding_bwd_workspace_bytes(int64_t n_tok, int64_t H);
int bf_embedding_bwd(const void* g, int32_t g_dtype, const int64_t* sorted_ids, const int64_t* perm, int64_t n_tok,
                     int64_t tok_per_sample, const float* rho, int64_t V, int64_t H, int64_t padding_idx,
                     uint64_t seed, uint32_t step, uint32_t tensor_id, const float* eps_in, float* grad_mu,
                     float* grad_rho, void* workspace, void* stream);

/* ------------------------------------------------------------------------- *
 * Fused  y = LayerNorm(dropout(h) + r)  around a Bayesian Linear (the "output"
 * blocks of a transformer host model: dense -> dropout -> LayerNorm(h + input),
 * e.g. transformers' BertSelfOutput / BertOutput).  The LayerNorm is the S-sample
 * LayerNorm above (per-sample affine) or the host model's
 * frequentist one (affine_stride == 0).
 *
 *   fwd : z = dropout_p(h) + r (kept for bwd, rounded to `dtype`), y = LN(z)*gamma_s + beta_s
 *   bwd : dz (gradient of r), dh = dz * keep / (1 - p) (gradient of h),
 *         dgamma, dbeta, and dbias[s][:] = sum_m dh[s][m][:]  -- the bias gradient of the
 *         Linear that produced h, so that layer skips its bf_bias_grad pass
 *
 * The keep mask is never stored: keep(e) = u16(e) >= round(p * 65536), where the 8 u16 of
 * elements [8k, 8k+8) are the 16-bit halves (low half first) of the four words of
 * Philox4x32-10(counter = (k & 0xffffffff, k >> 32, 0x80000000 | site_id, step [+ device
 * step counter]), key = seed); backward regenerates it.  bf_dropout_mask writes that
 * mask as bytes (for tests).  p_drop == 0 switches dropout off (dh may be NULL; dh == dz).
 *
 * h, r, z, y, gy, dz, dh  [S*M, H] of `dtype`; gamma, beta, mean, rstd as for bf_layernorm_*
 * dbias      [S, H] fp32 or NULL;  dgamma / dbeta  [S, H] (affine_stride != 0) or [H]
 * workspace  bf_resln_bwd_workspace_bytes(S, M, H) bytes, zero-filled once
 * ------------------------------------------------------------------------- */
int bf_resln_supported(int64_t H);
int bf_resln_fwd(const void* h, const void* r, int32_t dtype, const float* gamma, const float* beta,
                 int64_t affine_stride, int64_t S, int64_t M, int64_t H, float eps, float p_drop, uint64_t seed,
                 uint32_t step, uint32_t site_id, void* z, void* y, float* mean, float* rstd, void* stream);
int64_t bf_resln_bwd_workspace_bytes(int64_t S, int64_t M, int64_t H);
int bf_resln_bwd(const void* gy, const void* z, int32_t dtype, const float* gamma, int64_t affine_stride,
                 const float* mean, const float* rstd, int64_t S, int64_t M, int64_t H, float p_drop, uint64_t seed,
                 uint32_t step, uint32_t site_id, void* dz, void* dh, float* dgamma, float* dbeta, float* dbias,
                 void* workspace, void* stream);
/* the same with the keep bits handed from forward to backward instead of regenerated: keep [S*M, 32] uint32 -- word l
 * of a row holds the keep bytes of the octets c*32 + l (byte c, c < H/256; bit j of a byte = element j of the octet kept)
 * -- written by _fwd_keep, read by _bwd_keep: the backward is issue-bound and the Philox regeneration is half of its
 * instructions; costs 128 bytes per row each way (one coalesced store / load per row).  keep == NULL: the plain functions. */
int bf_resln_fwd_keep(const void* h, const void* r, int32_t dtype, const float* gamma, const float* beta,
                      int64_t affine_stride, int64_t S, int64_t M, int64_t H, float eps, float p_drop, uint64_t seed,
                      uint32_t step, uint32_t site_id, void* z, void* y, float* mean, float* rstd, uint32_t* keep,
                      void* stream);
int bf_resln_bwd_keep(const void* gy, const void* z, int32_t dtype, const float* gamma, int64_t affine_stride,
                      const float* mean, const float* rstd, int64_t S, int64_t M, int64_t H, float p_drop, uint64_t seed,
                      uint32_t step, uint32_t site_id, void* dz, void* dh, float* dgamma, float* dbeta, float* dbias,
                      void* workspace, const uint32_t* keep, void* stream);
int bf_dropout_mask(uint8_t* out, int64_t n, float p_drop, uint64_t seed, uint32_t step, uint32_t site_id,
                    void* stream);

/* ------------------------------------------------------------------------- *
 * Short-sequence self-attention of the host model between the Bayesian q / k / v and output projections
 * (opt-in extension; the reference leaves attention to the host model, i.e. to torch's
 * scaled_dot_product_attention):  O = dropout_p(softmax(q k^T * scale)) v  for whole sequences of
 * T <= 128 tokens (T % 16 == 0) and head width 64, one (sequence, head) per thread block, forward
 * and backward; no attention mask, not causal.  bf16 in / out, fp32 accumulation and softmax.
 *
 * q, k, v   bf16, element (b, h, t, j) at  base + b*strides[3i] + h*strides[3i+1] + t*strides[3i+2] + j
 *           (i = 0 q, 1 k, 2 v; unit inner stride): read in place from the [B, T, heads*64] projection
 *           outputs, no transposed copies
 * out, dout, dq, dk, dv   bf16 [B, T, H, 64] densely packed
 * lse       fp32 [B, H, T]: base-2 log-sum-exp of the scaled scores, written by fwd, read by bwd
 * keep      optional uint32 [B, H, T, 4] (only used when T == 128): the 128 keep bits of every query row (bit k % 32
 *           of word k / 32), written by fwd and read by bwd so that the tcgen05 backward does not regenerate the
 *           mask.  Null: the mma.sync kernels run (they regenerate it), whatever BF_OPT_ATTN_TC says, unless
 *           p_drop == 0.
 * out       (bwd) the forward output; unused since round 2 (D = rowsum(dO o O) is formed from P and dP), may be null
 * The dropout keep mask is a pure function of (seed, site, step [+ device step counter], b, h, q, k)
 * (Philox4x32-10; see bf_attention.cu), the same for both kernel families; bf_attention_dropout_mask
 * writes it as bytes [B, H, T, T] (tests).  Deterministic: every output element is written once.
 * T == 128 runs on the tensor cores' tcgen05 path (bf_attention_tc.cu: TMA-loaded tiles, TMEM accumulators, one row
 * per thread for the softmax); other lengths use mma.sync on ldmatrix fragments (bf_attention.cu).
 * ------------------------------------------------------------------------- */
int bf_attention_supported(int64_t T, int64_t head_dim);
int bf_attention_fwd(const void* q, const void* k, const void* v, const int64_t* strides, int64_t B, int64_t H, int64_t T,
                     float scale, float p_drop, uint64_t seed, uint32_t step, uint32_t site, void* out, float* lse,
                     uint32_t* keep, void* stream);
int bf_attention_bwd(const void* dout, const void* q, const void* k, const void* v, const int64_t* strides,
                     const void* out, const float* lse, const uint32_t* keep, int64_t B, int64_t H, int64_t T, float scale,
                     float p_drop, uint64_t seed, uint32_t step, uint32_t site, void* dq, void* dk, void* dv, void* stream);
/* bf_attention_bwd that also emits the bias gradients of the q / k / v projections, i.e. the column sums of dq, dk, dv
 * over the rows of each of the S folded samples (B % S == 0; sample s = sequences [s*B/S, (s+1)*B/S)):
 *   dbias  fp32 [3][S][H*64] (q, k, v), taken from the gradient tiles staged in shared memory (per-block partial rows in
 *   `workspace`, added in block order by a second small pass: deterministic).  T == 128 / tcgen05 path only. */
int64_t bf_attention_bias_workspace_bytes(int64_t B, int64_t H, int64_t S);
int bf_attention_bwd_bias(const void* dout, const void* q, const void* k, const void* v, const int64_t* strides,
                          const void* out, const float* lse, const uint32_t* keep, int64_t B, int64_t H, int64_t T,
                          float scale, float p_drop, uint64_t seed, uint32_t step, uint32_t site, void* dq, void* dk,
                          void* dv, float* dbias, void* workspace, int64_t S, void* stream);
int bf_attention_dropout_mask(uint8_t* out, int64_t B, int64_t H, int64_t T, float p_drop, uint64_t seed, uint32_t step,
                              uint32_t site, void* stream);

/* ------------------------------------------------------------------------- *
 * Fused global-norm clipping + AdamW over all trainable tensors.
 * Replaces clip_grad_norm_(params, max_norm) + AdamW.step() of the reference's
 * training loop (examples/bert_glue.py:240-241); arithmetic of torch.optim.AdamW
 * (decoupled weight decay, bias correction).
 *
 * descs      DEVICE array of per-tensor descriptors; `grad == NULL` skips a tensor
 * chunks     DEVICE int32 pairs {tensor index, chunk index}; a chunk is
 *            bf_optim_chunk_elems() consecutive elements of one tensor
 * steps      DEVICE float[n_tensors], zero-initialised by the caller: per-tensor update counts
 *            (torch semantics: a tensor without a gradient does not advance); incremented by the
 *            call itself, device-resident so a captured CUDA graph advances them on replay
 * max_norm   <= 0 disables clipping
 * grad_norm_out  NULL or DEVICE float[1]: global L2 norm of the unclipped gradients
 * workspace  bf_clip_adamw_workspace_bytes(n_chunks) bytes, contents irrelevant
 * ------------------------------------------------------------------------- */
typedef struct bf_opt_desc {
    void* param;       /* [n] of dtype, updated in place */
    const void* grad;  /* [n] of dtype */
    float* exp_avg;    /* [n] fp32 */
    float* exp_avg_sq; /* [n] fp32 */
    int64_t n;
    int32_t dtype; /* BF_F32 / BF_BF16 (param and grad) */
    int32_t vec;   /* 1: n % 4 == 0 and all pointers 16 B aligned (8 B for bf16 param / grad) */
    float* master; /* NULL, or [n] fp32 master copy of a BF_BF16 parameter: the update reads and writes the master
                      and stores its bf16 rounding in `param` (an update smaller than half a bf16 ulp is not lost) */
    float* sigma_out; /* NULL, or [n] fp32: for a rho tensor, softplus(updated rho) is written next to the update, so
                         the next forward's sampling reads sigma instead of recomputing it (bf_tensor_desc.sigma) */
} bf_opt_desc;

int32_t bf_optim_chunk_elems(void);
int64_t bf_clip_adamw_workspace_bytes(int64_t n_chunks);
int bf_clip_adamw_step(const bf_opt_desc* descs, const int32_t* chunks, int32_t n_chunks, float lr, float beta1,
                       float beta2, float eps, float weight_decay, float max_norm, float* steps,
                       float* grad_norm_out, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BAYEFORMERS_B200_H_ */
