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
const uint32_t* g_step_counter = nullptr;
}
const uint32_t* bf_step_counter() { return g_step_counter; }

extern "C" int bf_set_step_counter(const uint32_t* device_counter) {
    g_step_counter = device_counter;
    return 0;
}

extern "C" int bf_abi_version(void) { return BF_ABI_VERSION; }

extern "C" const char* bf_last_error(void) { return g_last_error.c_str(); }

extern "C" int bf_device_is_sm100(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
    return major == 10 ? 1 : 0;
}
