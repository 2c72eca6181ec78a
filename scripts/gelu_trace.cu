// Phase timeline of the fused bias + GELU forward contraction (cluster 0's leader block, first 16 tiles): builds
// bf_gemm_act.cu with -DBF_GELU_TRACE into a standalone binary and prints clock deltas.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DBF_GELU_TRACE \
//        scripts/gelu_trace.cu -o scripts/_bin/gelu_trace -Lbayeformers_b200 -l:libbayeformers_b200.so -lcuda
#include "../bayeformers_b200/csrc/bf_gemm_act.cu"

#include <cstdio>
#include <vector>

int main() {
    const int64_t S = 4, M = 65536, N = 3072, K = 768;
    __nv_bfloat16 *x, *w, *z, *y;
    float* bias;
    cudaMalloc(&x, S * M * K * 2), cudaMalloc(&w, S * N * K * 2), cudaMalloc(&z, S * M * N * 2), cudaMalloc(&y, S * M * N * 2);
    cudaMalloc(&bias, S * N * 4);
    std::vector<__nv_bfloat16> h((size_t)S * M * K);
    for (size_t i = 0; i < h.size(); ++i) h[i] = __float2bfloat16(((int)((i * 2654435761u) >> 20 & 1023) - 512) / 512.0f);
    cudaMemcpy(x, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(w, h.data(), (size_t)S * N * K * 2, cudaMemcpyHostToDevice);
    cudaMemset(bias, 0, S * N * 4);
    for (int rep = 0; rep < 3; ++rep) {
        const int rc = bf_linear_fwd_gelu(x, w, bias, z, y, S, M, N, K, 0);
        if (rc) { printf("rc %d\n", rc); return 1; }
    }
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda: %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long t[3 * 16 * 16];
    cudaMemcpyFromSymbol(t, act::g_gelu_trace, sizeof(t));
    printf("MMA thread: tile period, wait for a free accumulator, issue of the 12 k-steps\n");
    for (int i = 2; i < 12; ++i)
        printf("tile %2d: period %6llu  acc wait %5llu  issue %6llu\n", i, t[i * 16] - t[(i - 1) * 16], t[i * 16 + 1] - t[i * 16],
               t[i * 16 + 2] - t[i * 16 + 1]);
    for (int g = 0; g < 2; ++g) {
        printf("epilogue group %d (store thread): per tile [wait tfull], then per half box: ld | gelu + staging | wait for the other set's store\n", g);
        const unsigned long long* e0 = t + (1 + g) * 256;
        for (int i = 2; i < 10; ++i) {
            printf("tile %2d: period %6llu tfull wait %5llu", i, e0[i * 16] - e0[(i - 1) * 16], e0[i * 16 + 1] - e0[i * 16]);
            unsigned long long prev = e0[i * 16 + 1];
            for (int k = 0; k < 4; ++k) {
                printf(" | hb%d: ld(+bar) %5llu, compute %5llu, store wait %5llu", k, e0[i * 16 + 2 + 3 * k] - prev,
                       e0[i * 16 + 3 + 3 * k] - e0[i * 16 + 2 + 3 * k], e0[i * 16 + 4 + 3 * k] - e0[i * 16 + 3 + 3 * k]);
                prev = e0[i * 16 + 4 + 3 * k];
            }
            printf(" | to next tile top %5llu\n", e0[(i + 1) * 16] - prev);
        }
    }
    return 0;
}
