// RICK Fisher-information step on the device (sm_100a): grad^2 accumulation, per-filter FIM, percentile
// thresholds, freeze/prune decisions and the per-iteration mask application.
//
// Replaces host-side NumPy in train_dynamic_update_prune.py (see include/rick_b200.h for the line map).
// Everything here is order-sensitive on purpose: masks must be BIT-EXACT against the reference given the same
// Fisher tensor, so
//   * accumulation rounds g*g to float32 before the add (no FMA contraction),
//   * the per-filter mean reproduces NumPy's float32 pairwise summation tree (leaf blocks <= 128 elements with
//     8 interleaved partial sums, halves split at a multiple of 8) and then divides in float32,
//   * thresholds are exact order statistics (radix select on the float bit patterns) combined with NumPy's
//     'linear' lerp in float64, and comparisons are done in float64.
// All kernels are HBM- or latency-bound integer / float32 streaming work; no tensor cores involved.
#include "common.cuh"

namespace rick {

// ------------------------------------------------------------------------------------------ multi-tensor tables
constexpr int kMaxTensors = 48;          // per launch; host loops for longer lists (kernel parameter space is 4 KB)
constexpr int kChunk = 256 * 16;         // elements per CTA work item

struct AccumTable {
    float* acc[kMaxTensors];
    const float* grad[kMaxTensors];
    long long numel[kMaxTensors];
    int block_end[kMaxTensors];          // inclusive prefix sum of CTAs per tensor
    int count;
};

__device__ __forceinline__ int find_tensor(const int* block_end, int count, int b) {
    int t = 0;
    while (t < count - 1 && b >= block_end[t]) ++t;
    return t;
}

template <int MODE>  // 0: acc = g*g   1: acc += g*g   2: acc /= divisor
__global__ void __launch_bounds__(256) fisher_multi(const __grid_constant__ AccumTable tab, float divisor) {
    const int t = find_tensor(tab.block_end, tab.count, blockIdx.x);
    const int local_block = blockIdx.x - (t ? tab.block_end[t - 1] : 0);
    float* __restrict__ acc = tab.acc[t];
    const float* __restrict__ g = tab.grad[t];
    const long long n = tab.numel[t];
    const long long begin = (long long)local_block * kChunk;
    const long long end = min(begin + (long long)kChunk, n);
    const bool vec = ((reinterpret_cast<uintptr_t>(acc) | (MODE == 2 ? 0 : reinterpret_cast<uintptr_t>(g))) & 15) == 0;
    if (vec) {
        const long long vend = begin + ((end - begin) & ~3LL);
        for (long long i = begin + threadIdx.x * 4; i < vend; i += 256 * 4) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), x = a;
            if (MODE != 0) a = *reinterpret_cast<const float4*>(acc + i);
            if (MODE != 2) x = ld_stream_f4(reinterpret_cast<const float4*>(g + i));
            if (MODE == 2) {
                a.x = __fdiv_rn(a.x, divisor), a.y = __fdiv_rn(a.y, divisor);
                a.z = __fdiv_rn(a.z, divisor), a.w = __fdiv_rn(a.w, divisor);
            } else {
                a.x = __fadd_rn(a.x, __fmul_rn(x.x, x.x)), a.y = __fadd_rn(a.y, __fmul_rn(x.y, x.y));
                a.z = __fadd_rn(a.z, __fmul_rn(x.z, x.z)), a.w = __fadd_rn(a.w, __fmul_rn(x.w, x.w));
            }
            *reinterpret_cast<float4*>(acc + i) = a;
        }
        for (long long i = vend + threadIdx.x; i < end; i += 256) {
            if (MODE == 2) acc[i] = __fdiv_rn(acc[i], divisor);
            else acc[i] = __fadd_rn(MODE ? acc[i] : 0.f, __fmul_rn(g[i], g[i]));
        }
    } else {
        for (long long i = begin + threadIdx.x; i < end; i += 256) {
            if (MODE == 2) acc[i] = __fdiv_rn(acc[i], divisor);
            else acc[i] = __fadd_rn(MODE ? acc[i] : 0.f, __fmul_rn(g[i], g[i]));
        }
    }
}

static int run_multi(int mode, float* const* acc, const float* const* grad, const int64_t* numel, int count,
                     float divisor, cudaStream_t s) {
    for (int base = 0; base < count; base += kMaxTensors) {
        AccumTable tab{};
        const int m = (count - base < kMaxTensors) ? count - base : kMaxTensors;
        long long blocks = 0;
        int used = 0;
        for (int i = 0; i < m; ++i) {
            if (numel[base + i] < 0 || !acc[base + i] || (mode != 2 && !grad[base + i])) return RICK_ERR_INVALID_ARGUMENT;
            if (numel[base + i] == 0) continue;
            tab.acc[used] = acc[base + i];
            tab.grad[used] = mode != 2 ? grad[base + i] : nullptr;
            tab.numel[used] = numel[base + i];
            blocks += ceil_div(numel[base + i], kChunk);
            if (blocks > 0x7fffffffLL) return RICK_ERR_OVERFLOW;
            tab.block_end[used] = (int)blocks;
            ++used;
        }
        if (!used) continue;
        tab.count = used;
        if (mode == 0) fisher_multi<0><<<(unsigned)blocks, 256, 0, s>>>(tab, divisor);
        else if (mode == 1) fisher_multi<1><<<(unsigned)blocks, 256, 0, s>>>(tab, divisor);
        else fisher_multi<2><<<(unsigned)blocks, 256, 0, s>>>(tab, divisor);
        RICK_CHECK_LAUNCH();
    }
    return RICK_OK;
}

// ------------------------------------------------------------------------------------------ per-filter FIM
// The pairwise-summation tree for a row of `len` elements is planned on the host: leaves in order, plus for each
// leaf the number of (pop two, push sum) reductions that follow it -- a postfix encoding of NumPy's recursion.
constexpr int kMaxLeaves = 448;
struct SumPlan {
    int len;
    int nleaves;
    unsigned short leaf_size[kMaxLeaves];   // <= 128
    unsigned char reduces[kMaxLeaves];
};

static void plan_rec(SumPlan& pl, int n, bool& ok) {
    if (!ok) return;
    if (n <= 128) {
        if (pl.nleaves >= kMaxLeaves) { ok = false; return; }
        pl.leaf_size[pl.nleaves] = (unsigned short)n;
        pl.reduces[pl.nleaves] = 0;
        ++pl.nleaves;
        return;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    plan_rec(pl, n2, ok);
    plan_rec(pl, n - n2, ok);
    if (ok) ++pl.reduces[pl.nleaves - 1];
}

__device__ __forceinline__ float leaf_sum(const float* __restrict__ a, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, __ldg(a + i));
        return res;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __ldg(a + j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], __ldg(a + i + j));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, __ldg(a + i));
    return res;
}

// one warp per filter row
__global__ void __launch_bounds__(128) filter_fim_kernel(float* __restrict__ fim, const float* __restrict__ w,
                                                         const float* __restrict__ b, long long rows,
                                                         const __grid_constant__ SumPlan plan) {
    __shared__ float s_leaf[4][kMaxLeaves];
    __shared__ int s_off[kMaxLeaves];
    for (int l = threadIdx.x; l < plan.nleaves; l += blockDim.x) {
        int off = 0;                       // small plans: a serial prefix per leaf is fine
        for (int k = 0; k < l; ++k) off += plan.leaf_size[k];
        s_off[l] = off;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long row = blockIdx.x * 4LL + warp; row < rows; row += gridDim.x * 4LL) {
        const float* a = w + row * plan.len;
        for (int l = lane; l < plan.nleaves; l += 32) s_leaf[warp][l] = leaf_sum(a + s_off[l], plan.leaf_size[l]);
        __syncwarp();
        if (lane == 0) {
            float stack[24];
            int sp = 0;
            for (int l = 0; l < plan.nleaves; ++l) {
                stack[sp++] = s_leaf[warp][l];
                for (int r = plan.reduces[l]; r > 0; --r) {
                    stack[sp - 2] = __fadd_rn(stack[sp - 2], stack[sp - 1]);
                    --sp;
                }
            }
            float m = __fdiv_rn(stack[0], (float)plan.len);
            if (b) m = __fdiv_rn(__fadd_rn(m, __ldg(b + row)), 2.0f);
            fim[row] = m;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------ percentile
__device__ __forceinline__ unsigned key_of(float f) {   // order-preserving float -> uint
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_of(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct QuantileJobs {
    long long lo[8];     // rank of the lower order statistic
    long long hi[8];     // rank of the upper one (== lo at the top end)
    double gamma[8];
    int nq;
};

// single CTA; n is small (thousands of filters).  MSB-first 8-bit radix select, one job after another.
__global__ void __launch_bounds__(1024) percentile_kernel(double* __restrict__ lines, const float* __restrict__ v,
                                                          long long n, const __grid_constant__ QuantileJobs jobs) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_mask;
    __shared__ long long s_rank;
    __shared__ unsigned s_min_above;
    __shared__ unsigned long long s_count_le;
    for (int j = 0; j < jobs.nq; ++j) {
        if (threadIdx.x == 0) s_prefix = 0u, s_mask = 0u, s_rank = jobs.lo[j];
        __syncthreads();
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
            __syncthreads();
            const unsigned prefix = s_prefix, mask = s_mask;
            for (long long i = threadIdx.x; i < n; i += blockDim.x) {
                const unsigned k = key_of(__ldg(v + i));
                if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                long long r = s_rank;
                unsigned d = 0;
                for (; d < 256; ++d) {
                    if (r < (long long)hist[d]) break;
                    r -= hist[d];
                }
                s_rank = r;
                s_prefix = prefix | (d << shift);
                s_mask = mask | (255u << shift);
            }
            __syncthreads();
        }
        const unsigned klo = s_prefix;                 // key of the lo-th smallest element
        if (threadIdx.x == 0) s_min_above = 0xffffffffu, s_count_le = 0ull;
        __syncthreads();
        unsigned my_min = 0xffffffffu;
        unsigned my_le = 0;
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned k = key_of(__ldg(v + i));
            if (k <= klo) ++my_le;
            else my_min = min(my_min, k);
        }
        atomicMin(&s_min_above, my_min);
        atomicAdd(&s_count_le, (unsigned long long)my_le);
        __syncthreads();
        if (threadIdx.x == 0) {
            const double a = (double)float_of(klo);
            double b = a;
            if (jobs.hi[j] != jobs.lo[j] && (long long)s_count_le <= jobs.hi[j]) b = (double)float_of(s_min_above);
            const double diff = __dsub_rn(b, a);
            const double g = jobs.gamma[j];
            lines[j] = (g >= 0.5) ? __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, g))) : __dadd_rn(a, __dmul_rn(diff, g));
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------ decisions
__global__ void __launch_bounds__(256) decide_kernel(uint8_t* __restrict__ state, uint8_t* __restrict__ zero,
                                                     const float* __restrict__ fim, long long n,
                                                     const double* __restrict__ lines, int flags, int reset_zero) {
    const bool closed_low = (flags & 1) != 0;
    // NumPy >= 2 compares a float32 array with a float64 scalar in float64; NumPy < 2 (the reference pins 1.23.1,
    // environment.yml:49) casts the scalar to float32 first (value-based casting).  The two differ only for a FIM
    // within half a float32 ulp of a threshold.
    const bool f32_compare = (flags & RICK_DECIDE_COMPARE_F32) != 0;
    const double cut = f32_compare ? (double)(float)lines[0] : lines[0];
    const double pr = f32_compare ? (double)(float)lines[1] : lines[1];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double f = (double)fim[i];
        uint8_t s = 0;
        if (f > cut) s |= 1;
        const bool prune = closed_low ? (f < pr) : (f <= pr);
        if (prune) s |= 2;
        if ((closed_low ? (f >= pr) : (f > pr)) && f <= cut) s |= 4;
        state[i] = s;
        if (zero) zero[i] = (uint8_t)(((reset_zero ? 0 : zero[i]) | (prune ? 1 : 0)) ? 1 : 0);
    }
}

// ------------------------------------------------------------------------------------------ mask application
constexpr int kMaxMaskTensors = 64;
struct MaskTable {
    float* param[kMaxMaskTensors];
    float* grad[kMaxMaskTensors];
    const uint8_t* state[kMaxMaskTensors];
    const uint8_t* zero[kMaxMaskTensors];
    long long rows[kMaxMaskTensors];
    long long inner[kMaxMaskTensors];
    int block_end[kMaxMaskTensors];
    int count;
};
constexpr int kMaskChunk = 256 * 8;

__global__ void __launch_bounds__(256) mask_apply_kernel(const __grid_constant__ MaskTable tab) {
    const int t = find_tensor(tab.block_end, tab.count, blockIdx.x);
    const int local_block = blockIdx.x - (t ? tab.block_end[t - 1] : 0);
    float* __restrict__ param = tab.param[t];
    float* __restrict__ grad = tab.grad[t];
    const uint8_t* __restrict__ state = tab.state[t];
    const uint8_t* __restrict__ zero = tab.zero[t];
    const long long inner = tab.inner[t], total = tab.rows[t] * inner;
    const long long begin = (long long)local_block * kMaskChunk;
    const long long end = min(begin + (long long)kMaskChunk, total);
    for (long long i = begin + threadIdx.x; i < end; i += 256) {
        const long long row = i / inner;
        const bool z = zero && zero[row];
        const bool f = state && (state[row] & 1);
        if (z && param) param[i] = 0.f;
        if ((z || f) && grad) grad[i] = 0.f;
    }
}

}  // namespace rick

extern "C" int rick_fisher_accum(float* const* acc, const float* const* grad, const int64_t* numel, int count,
                                 int first, rick_stream_t stream) {
    if (!acc || !grad || !numel || count < 0) return RICK_ERR_INVALID_ARGUMENT;
    return rick::run_multi(first ? 0 : 1, acc, grad, numel, count, 1.f, static_cast<cudaStream_t>(stream));
}

extern "C" int rick_fisher_divide(float* const* acc, const int64_t* numel, int count, float divisor,
                                  rick_stream_t stream) {
    if (!acc || !numel || count < 0) return RICK_ERR_INVALID_ARGUMENT;
    return rick::run_multi(2, acc, nullptr, numel, count, divisor, static_cast<cudaStream_t>(stream));
}

extern "C" int rick_filter_fim(float* fim, const float* fisher_w, const float* fisher_b, int64_t rows, int64_t len,
                               rick_stream_t stream) {
    using namespace rick;
    if (!fim || !fisher_w || rows < 0 || len < 1) return RICK_ERR_INVALID_ARGUMENT;
    if (len > 0x7fffffff) return RICK_ERR_OVERFLOW;
    if (rows == 0) return RICK_OK;
    SumPlan plan{};
    plan.len = (int)len;
    bool ok = true;
    plan_rec(plan, (int)len, ok);
    if (!ok) return RICK_ERR_UNSUPPORTED;   // rows longer than ~28k elements: not a filter shape of this model
    long long blocks = ceil_div(rows, 4);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    filter_fim_kernel<<<(unsigned)blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(fim, fisher_w, fisher_b, rows,
                                                                                         plan);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

extern "C" int rick_percentile(double* lines, const float* fim, int64_t n, const double* q, int nq,
                               rick_stream_t stream) {
    using namespace rick;
    if (!lines || !fim || !q || n < 1 || nq < 1 || nq > 8) return RICK_ERR_INVALID_ARGUMENT;
    QuantileJobs jobs{};
    jobs.nq = nq;
    for (int j = 0; j < nq; ++j) {
        if (!(q[j] >= 0.0 && q[j] <= 100.0)) return RICK_ERR_INVALID_ARGUMENT;
        // numpy/lib/_function_base_impl.py: q/100, virtual index (n-1)*q, floor, gamma = virtual - floor
        const double quant = q[j] / 100.0;
        const double virt = (double)(n - 1) * quant;
        double prev = floor(virt);
        long long lo = (long long)prev, hi = lo + 1;
        double gamma = virt - prev;
        if (virt >= (double)(n - 1)) lo = hi = n - 1;   // indexes above bounds -> last element for both
        if (lo < 0) lo = 0;
        if (hi > n - 1) hi = n - 1;
        jobs.lo[j] = lo, jobs.hi[j] = hi, jobs.gamma[j] = gamma;
    }
    percentile_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(lines, fim, n, jobs);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

extern "C" int rick_decide(uint8_t* state, uint8_t* zero_mask, const float* fim, int64_t n, const double* lines,
                           int flags, int reset_zero, rick_stream_t stream) {
    using namespace rick;
    if (!state || !fim || !lines || n < 0) return RICK_ERR_INVALID_ARGUMENT;
    if (n == 0) return RICK_OK;
    long long blocks = ceil_div(n, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    decide_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(state, zero_mask, fim, n, lines,
                                                                                    flags, reset_zero);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

extern "C" int rick_mask_apply(float* const* param, float* const* grad, const uint8_t* const* state,
                               const uint8_t* const* zero, const int64_t* rows, const int64_t* inner, int count,
                               rick_stream_t stream) {
    using namespace rick;
    if (!param || !grad || !state || !zero || !rows || !inner || count < 0) return RICK_ERR_INVALID_ARGUMENT;
    for (int base = 0; base < count; base += kMaxMaskTensors) {
        MaskTable tab{};
        const int m = (count - base < kMaxMaskTensors) ? count - base : kMaxMaskTensors;
        long long blocks = 0;
        int used = 0;
        for (int i = 0; i < m; ++i) {
            const int k = base + i;
            if (rows[k] < 0 || inner[k] < 1) return RICK_ERR_INVALID_ARGUMENT;
            if (rows[k] == 0 || (!state[k] && !zero[k]) || (!param[k] && !grad[k])) continue;
            tab.param[used] = param[k], tab.grad[used] = grad[k];
            tab.state[used] = state[k], tab.zero[used] = zero[k];
            tab.rows[used] = rows[k], tab.inner[used] = inner[k];
            blocks += ceil_div(rows[k] * inner[k], kMaskChunk);
            if (blocks > 0x7fffffffLL) return RICK_ERR_OVERFLOW;
            tab.block_end[used] = (int)blocks;
            ++used;
        }
        if (!used) continue;
        tab.count = used;
        mask_apply_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(tab);
        RICK_CHECK_LAUNCH();
    }
    return RICK_OK;
}
