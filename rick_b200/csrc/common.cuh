// Shared helpers for the rick_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rick_b200.h"

namespace rick {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids below are sized in multiples of this

void set_cuda_error(cudaError_t e);
void count_launch();   // statistics only: number of kernels this library has launched (rick_launch_count)

// Call right after a kernel launch: turns a launch failure into RICK_ERR_CUDA.
#define RICK_CHECK_LAUNCH()                               \
    do {                                                  \
        cudaError_t e__ = cudaGetLastError();             \
        if (e__ != cudaSuccess) {                         \
            ::rick::set_cuda_error(e__);                  \
            return RICK_ERR_CUDA;                         \
        }                                                 \
        ::rick::count_launch();                           \
    } while (0)

#define RICK_CUDA_TRY(expr)                               \
    do {                                                  \
        cudaError_t e__ = (expr);                         \
        if (e__ != cudaSuccess) {                         \
            ::rick::set_cuda_error(e__);                  \
            return RICK_ERR_CUDA;                         \
        }                                                 \
    } while (0)

static inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int floor_div(int a, int b) {
    int q = a / b;
    return ((a % b != 0) && ((a < 0) != (b < 0))) ? q - 1 : q;
}
static inline int floor_mod(int a, int b) { return a - floor_div(a, b) * b; }

// ---- element access that is uniform over fp32 / bf16 storage (arithmetic is always fp32) ----
template <typename T> struct Elem;
template <> struct Elem<float> {
    static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
    static __device__ __forceinline__ float ld(const __nv_bfloat16* p) {
        return __bfloat162float(__ldg(p));
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// streaming (read-once / write-once) 128-bit accesses: keep them out of L1
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u4(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// grad_bias[c] = sum_n sum_j src[(n*C + c) * len + j]; one warp per channel, fixed order -> deterministic
template <typename S>
__global__ void __launch_bounds__(256) bias_grad_fold(float* __restrict__ grad_bias, const S* __restrict__ src,
                                                      int n, int c, long long len) {
    const int lane = threadIdx.x & 31;
    const int ch = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    if (ch >= c) return;
    float acc = 0.f;
    for (int b = 0; b < n; ++b) {
        const S* row = src + ((long long)b * c + ch) * len;
        for (long long j = lane; j < len; j += 32) acc += Elem<S>::ld(row + j);
    }
    acc = warp_sum(acc);
    if (lane == 0) grad_bias[ch] = acc;
}


}  // namespace rick
