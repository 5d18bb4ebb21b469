// Status strings and per-thread CUDA error text for the rick_b200 C ABI.
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace rick {
static thread_local char g_cuda_error[256] = "";
void set_cuda_error(cudaError_t e) {
    const char* s = cudaGetErrorString(e);
    strncpy(g_cuda_error, s ? s : "unknown", sizeof(g_cuda_error) - 1);
    g_cuda_error[sizeof(g_cuda_error) - 1] = 0;
}
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace rick

extern "C" int rick_abi_version(void) { return RICK_B200_ABI_VERSION; }

extern "C" const char* rick_status_string(int status) {
    switch (status) {
        case RICK_OK: return "ok";
        case RICK_ERR_INVALID_ARGUMENT: return "invalid argument";
        case RICK_ERR_UNSUPPORTED: return "unsupported configuration";
        case RICK_ERR_OVERFLOW: return "size overflows kernel index type";
        case RICK_ERR_CUDA: return "CUDA error";
        case RICK_ERR_ALIGNMENT: return "misaligned pointer";
        default: return "unknown status";
    }
}

extern "C" const char* rick_last_cuda_error(void) { return rick::g_cuda_error; }

extern "C" unsigned long long rick_launch_count(void) { return rick::g_launches.load(std::memory_order_relaxed); }
