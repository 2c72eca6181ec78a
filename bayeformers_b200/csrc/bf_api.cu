// C-ABI plumbing: version, thread-local error text, device queries.
#include "bf_common.cuh"

namespace {
thread_local std::string g_last_error;
}

void bf_set_error(const std::string& msg) { g_last_error = msg; }

int bf_num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

namespace {
// one counter per device (a process may drive several devices; the pointer is only valid on the device it lives on)
const uint32_t* g_step_counter[64] = {nullptr};
// tuning switches, set explicitly through bf_set_option (no environment reads inside the library)
int32_t g_options[BF_OPT_COUNT] = {/*GEMM_2CTA*/ 1, /*WGRAD_2CTA*/ 1, /*RESLN_BWD_STAGED*/ 1, /*SK_PREFETCH*/ 1, /*ATTN_TC*/ 1, /*GELU_POLY*/ 1};
}
const uint32_t* bf_step_counter() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    return g_step_counter[dev];
}
int bf_option(int id) { return (id >= 0 && id < BF_OPT_COUNT) ? g_options[id] : 0; }

extern "C" int bf_set_step_counter(const uint32_t* device_counter) {
    int dev = 0;
    BF_CUDA_OK(cudaGetDevice(&dev));
    BF_CHECK_ARG(dev >= 0 && dev < 64, "device index out of range");
    g_step_counter[dev] = device_counter;
    return 0;
}

extern "C" int bf_set_option(int32_t option, int32_t value) {
    BF_CHECK_ARG(option >= 0 && option < BF_OPT_COUNT, "unknown option");
    g_options[option] = value;
    return 0;
}

extern "C" int bf_get_option(int32_t option) { return bf_option(option); }

extern "C" int bf_abi_version(void) { return BF_ABI_VERSION; }

extern "C" const char* bf_last_error(void) { return g_last_error.c_str(); }

extern "C" int bf_device_is_sm100(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
    return major == 10 ? 1 : 0;
}
