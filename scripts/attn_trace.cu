// Phase timeline of the tcgen05 attention backward (block 0, first 16 pairs): builds bf_attention_tc.cu with
// -DBF_ATTN_TRACE into a standalone binary, runs it at the bench shape and prints clock deltas per phase.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DBF_ATTN_TRACE \
//        scripts/attn_trace.cu -o scripts/_bin/attn_trace -Lbayeformers_b200 -l:libbayeformers_b200.so -lcuda
#include "../bayeformers_b200/csrc/bf_attention_tc.cu"

#include <cstdio>
#include <cstdlib>
#include <vector>

int main() {
    const int64_t B = 2048, H = 12, T = 128, D = 64;
    const size_t n = (size_t)B * T * H * D;
    __nv_bfloat16 *q, *k, *v, *dO, *dq, *dk, *dv, *o;
    float* lse;
    uint32_t* keep;
    for (auto pp : {&q, &k, &v, &dO, &dq, &dk, &dv, &o}) cudaMalloc(pp, n * 2);
    cudaMalloc(&lse, (size_t)B * H * T * 4);
    cudaMalloc(&keep, (size_t)B * H * T * 16);
    std::vector<__nv_bfloat16> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16(((int)((i * 2654435761u) >> 20 & 1023) - 512) / 1024.0f);
    for (auto pp : {q, k, v, dO}) cudaMemcpy(pp, h.data(), n * 2, cudaMemcpyHostToDevice);
    const int64_t strides[9] = {T * H * D, D, H * D, T * H * D, D, H * D, T * H * D, D, H * D};
    for (int rep = 0; rep < 3; ++rep) {
        int rc = bf_attention_tc_fwd(q, k, v, strides, B, H, 0.125f, 0.1f, 1, 2, 3, o, lse, keep, 0);
        rc |= bf_attention_tc_bwd(dO, q, k, v, strides, lse, keep, B, H, 0.125f, 0.1f, 1, 2, 3, dq, dk, dv, nullptr, nullptr, 1, 0);
        if (rc) { printf("rc %d\n", rc); return 1; }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda: %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long t[2 * 16 * 16];
    if (getenv("TRACE_FWD")) {
        cudaMemset(0, 0, 0);
        bf_attention_tc_fwd(q, k, v, strides, B, H, 0.125f, 0.1f, 1, 2, 3, o, lse, keep, 0);
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(t, attn_tc::g_trace, sizeof(t));
        const char* fn[] = {"top", "s_ready", "max", "bar1", "bar1'", "exp+sum", "bar2", "select+staged P (then the drain of the previous pair)"};
        printf("forward, ALU thread 0 (cycles since the pair's top)\n");
        for (int i = 1; i < 12; ++i) {
            printf("pair %2d: top+%6llu", i, t[i * 16] - t[(i - 1) * 16]);
            for (int s = 1; s < 8; ++s) printf(" %s %llu", fn[s], t[i * 16 + s] - t[i * 16]);
            printf("\n");
        }
        return 0;
    }
    cudaMemcpyFromSymbol(t, attn_tc::g_trace, sizeof(t));
    const char* an[] = {"top", "bar1", "staged P,dS", "pass1(next)", "o_ready", "staged"};
    printf("ALU thread 0 (cycles since the pair's top; 'top' = since previous top)\n");
    for (int i = 1; i < 12; ++i) {
        printf("pair %2d:", i);
        printf(" top+%6llu", t[i * 16] - t[(i - 1) * 16]);
        for (int s = 1; s < 6; ++s) printf(" %s %llu", an[s], t[i * 16 + s] - t[i * 16]);
        printf("\n");
    }
    const char* mn[] = {"-", "-", "-", "S,dP(next) issued", "p_ready", "o_free", "outs issued"};
    printf("MMA thread (cycles since the ALU top of the same pair)\n");
    for (int i = 1; i < 12; ++i) {
        printf("pair %2d:", i);
        for (int s = 3; s < 7; ++s) printf(" %s %lld", mn[s], (long long)(t[(16 + i) * 16 + s] - t[i * 16]));
        printf("\n");
    }
    return 0;
}
