// Fused bias + leaky-ReLU * scale for sm_100a: forward, first derivative (with the per-channel bias-gradient
// reduction fused into the same pass) and second derivative.
//
// Replaces op/fused_bias_act_kernel.cu (fused_bias_act_op, :52-98) behind the same native signature
// (op/fused_bias_act.cpp:11-17), plus the ``grad_input.sum(dim)`` ATen reduction that op/fused_act.py:31-37 runs as
// a second full read of grad_input.  HBM-bound: forward moves 2 elements per output (read x, write out), the fused
// backward 3 (read grad_out, read saved out, write grad_in).
//
// Layout: flat contiguous tensor; the bias channel of element i is (i / step_b) % size_b with step_b = prod(dims[2:]).
// Vector path: 128-bit streaming loads/stores, one integer divide per 16 bytes (all lanes of a vector share a
// channel because step_b is a multiple of the vector width); 64-bit element indices.
#include "common.cuh"

namespace rick {

template <typename T> struct Vec16;  // 16 bytes of T <-> fp32 lanes
template <> struct Vec16<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        float4 t = ld_stream_f4(reinterpret_cast<const float4*>(p));
        v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        st_stream_f4(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    }
};
template <> struct Vec16<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        uint4 t = ld_stream_u4(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        st_stream_u4(reinterpret_cast<uint4*>(p), make_uint4(w[0], w[1], w[2], w[3]));
    }
};

__device__ __forceinline__ float act_apply(float v, float gate, int act, int grad, float alpha) {
    if (grad == 2) return 0.f;
    if (act == RICK_ACT_LRELU) return gate > 0.f ? v : v * alpha;
    return v;
}

// ------------------------------------------------------------------------------------------ forward / generic modes
template <typename T>
__global__ void __launch_bounds__(256) bias_act_vec(T* __restrict__ out, const T* __restrict__ x,
                                                    const T* __restrict__ bias, const T* __restrict__ ref,
                                                    long long nvec, long long step_vec, int size_b, int act, int grad,
                                                    float alpha, float scale) {
    constexpr int N = Vec16<T>::N;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float v[N], r[N];
        Vec16<T>::load(x + i * N, v);
        if (ref) Vec16<T>::load(ref + i * N, r);
        float b = 0.f;
        if (bias) b = Elem<T>::ld(bias + (int)((i / step_vec) % size_b));
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const float t = v[k] + b;
            v[k] = act_apply(t, (grad == 0) ? t : (ref ? r[k] : 0.f), act, grad, alpha) * scale;
        }
        Vec16<T>::store(out + i * N, v);
    }
}

// channels-last fp32: step_b == 1, the 4 lanes of a vector are 4 consecutive channels -> bias is a float4 at i % c4
__global__ void __launch_bounds__(256) bias_act_vec_nhwc(float* __restrict__ out, const float* __restrict__ x,
                                                         const float* __restrict__ bias, const float* __restrict__ ref,
                                                         long long nvec, int c4, int act, int grad, float alpha,
                                                         float scale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float4 v = ld_stream_f4(reinterpret_cast<const float4*>(x) + i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + (int)(i % c4));
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ref) r = ld_stream_f4(reinterpret_cast<const float4*>(ref) + i);
        v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
        v.x = act_apply(v.x, grad == 0 ? v.x : r.x, act, grad, alpha) * scale;
        v.y = act_apply(v.y, grad == 0 ? v.y : r.y, act, grad, alpha) * scale;
        v.z = act_apply(v.z, grad == 0 ? v.z : r.z, act, grad, alpha) * scale;
        v.w = act_apply(v.w, grad == 0 ? v.w : r.w, act, grad, alpha) * scale;
        st_stream_f4(reinterpret_cast<float4*>(out) + i, v);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) bias_act_scalar(T* __restrict__ out, const T* __restrict__ x,
                                                       const T* __restrict__ bias, const T* __restrict__ ref,
                                                       long long n, long long step_b, int size_b, int act, int grad,
                                                       float alpha, float scale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        float t = Elem<T>::ld(x + i);
        if (bias) t += Elem<T>::ld(bias + (int)((i / step_b) % size_b));
        const float gate = (grad == 0) ? t : (ref ? Elem<T>::ld(ref + i) : 0.f);
        Elem<T>::st(out + i, act_apply(t, gate, act, grad, alpha) * scale);
    }
}

template <typename T>
static int launch_bias_act(void* out, const void* x, const void* bias, const void* ref, int64_t n, int64_t step_b,
                           int64_t size_b, int act, int grad, float alpha, float scale, cudaStream_t s) {
    constexpr int N = Vec16<T>::N;
    const bool vec = (n % N == 0) && (!bias || step_b % N == 0) && aligned_to(out, 16) && aligned_to(x, 16) &&
                     (!ref || aligned_to(ref, 16));
    const long long cap = (long long)kNumSMs * 16;
    if (sizeof(T) == 4 && bias && step_b == 1 && size_b % 4 == 0 && n % 4 == 0 && aligned_to(out, 16) &&
        aligned_to(x, 16) && aligned_to(bias, 16) && (!ref || aligned_to(ref, 16))) {
        const long long nvec = n / 4;
        long long blocks = ceil_div(nvec, 256);
        if (blocks > cap) blocks = cap;
        bias_act_vec_nhwc<<<(unsigned)blocks, 256, 0, s>>>((float*)out, (const float*)x, (const float*)bias,
                                                            (const float*)ref, nvec, (int)(size_b / 4), act, grad, alpha,
                                                            scale);
        RICK_CHECK_LAUNCH();
        return RICK_OK;
    }
    if (vec) {
        const long long nvec = n / N;
        long long blocks = ceil_div(nvec, 256);
        if (blocks > cap) blocks = cap;
        bias_act_vec<T><<<(unsigned)blocks, 256, 0, s>>>((T*)out, (const T*)x, (const T*)bias, (const T*)ref, nvec,
                                                          bias ? step_b / N : 1, (int)size_b, act, grad, alpha, scale);
    } else {
        long long blocks = ceil_div(n, 256);
        if (blocks > cap) blocks = cap;
        bias_act_scalar<T><<<(unsigned)blocks, 256, 0, s>>>((T*)out, (const T*)x, (const T*)bias, (const T*)ref, n,
                                                             bias ? step_b : 1, (int)size_b, act, grad, alpha, scale);
    }
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

// ------------------------------------------------------------------------------------------ fused backward
// A warp owns one (plane, chunk) work item: chunk = 32 lanes x UNR vectors.  Partial sums go to
// workspace[plane * chunks + chunk]; a second small kernel folds them per channel in a fixed order.
constexpr int kBwdUnroll = 4;

template <typename T>
__global__ void __launch_bounds__(256) bias_act_bwd_main(T* __restrict__ gin, float* __restrict__ partial,
                                                         const T* __restrict__ gout, const T* __restrict__ saved,
                                                         long long planes, long long hw_vec, int chunks, float alpha,
                                                         float scale) {
    constexpr int N = Vec16<T>::N;
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long items = planes * chunks;
    for (long long item = warp0; item < items; item += nwarps) {
        const long long plane = item / chunks;
        const int chunk = (int)(item - plane * chunks);
        const long long base = plane * hw_vec;
        const long long v0 = (long long)chunk * (32 * kBwdUnroll);
        float g[kBwdUnroll][N], o[kBwdUnroll][N];
        bool ok[kBwdUnroll];
#pragma unroll
        for (int u = 0; u < kBwdUnroll; ++u) {
            const long long vi = v0 + u * 32 + lane;
            ok[u] = vi < hw_vec;
            if (ok[u]) {
                Vec16<T>::load(gout + (base + vi) * N, g[u]);
                Vec16<T>::load(saved + (base + vi) * N, o[u]);
            }
        }
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < kBwdUnroll; ++u) {
            if (!ok[u]) continue;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                g[u][k] = (o[u][k] > 0.f ? g[u][k] : g[u][k] * alpha) * scale;
                acc += g[u][k];
            }
            Vec16<T>::store(gin + (base + v0 + u * 32 + lane) * N, g[u]);
        }
        acc = warp_sum(acc);
        if (lane == 0) partial[item] = acc;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) bias_act_bwd_scalar(T* __restrict__ gin, const T* __restrict__ gout,
                                                           const T* __restrict__ saved, long long n, float alpha,
                                                           float scale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float g = Elem<T>::ld(gout + i);
        Elem<T>::st(gin + i, (Elem<T>::ld(saved + i) > 0.f ? g : g * alpha) * scale);
    }
}

// ---- channels-last: (rows, C) matrix, bias gradient = column sums.  A CTA owns kRowsPerCta rows; thread t owns the
// float4 column group t % c4 and walks rows t / c4, t / c4 + lanes, ...; row-lane partials are folded through shared
// memory and written as partial[c][cta] so that bias_grad_fold (fixed order) finishes the reduction.
constexpr int kRowsPerCta = 128;

__global__ void __launch_bounds__(256) bias_act_bwd_nhwc_main(float* __restrict__ gin, float* __restrict__ partial,
                                                              const float* __restrict__ gout,
                                                              const float* __restrict__ saved, long long rows, int c4,
                                                              int nctas, float alpha, float scale) {
    extern __shared__ float4 s_part[];            // (row_lanes, c4_tile)
    const int c4_tile = c4 < 256 ? c4 : 256;      // column groups handled per pass
    const int lanes = 256 / c4_tile;
    const int cg_l = threadIdx.x % c4_tile, rl = threadIdx.x / c4_tile;
    const long long row0 = (long long)blockIdx.x * kRowsPerCta;
    const long long row_end = min(row0 + (long long)kRowsPerCta, rows);
    for (int cg0 = 0; cg0 < c4; cg0 += c4_tile) {
        const int cg = cg0 + cg_l;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rl < lanes && cg < c4) {
            for (long long r = row0 + rl; r < row_end; r += lanes) {
                const long long i = r * c4 + cg;
                const float4 g = ld_stream_f4(reinterpret_cast<const float4*>(gout) + i);
                const float4 o = ld_stream_f4(reinterpret_cast<const float4*>(saved) + i);
                float4 v;
                v.x = (o.x > 0.f ? g.x : g.x * alpha) * scale, v.y = (o.y > 0.f ? g.y : g.y * alpha) * scale;
                v.z = (o.z > 0.f ? g.z : g.z * alpha) * scale, v.w = (o.w > 0.f ? g.w : g.w * alpha) * scale;
                st_stream_f4(reinterpret_cast<float4*>(gin) + i, v);
                acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
            }
        }
        if (rl < lanes) s_part[rl * c4_tile + cg_l] = acc;
        __syncthreads();
        if (rl == 0 && cg < c4) {
            float4 t = s_part[cg_l];
            for (int l = 1; l < lanes; ++l) {
                const float4 u = s_part[l * c4_tile + cg_l];
                t.x += u.x, t.y += u.y, t.z += u.z, t.w += u.w;
            }
            const long long c0 = (long long)cg * 4;
            partial[(c0 + 0) * nctas + blockIdx.x] = t.x, partial[(c0 + 1) * nctas + blockIdx.x] = t.y;
            partial[(c0 + 2) * nctas + blockIdx.x] = t.z, partial[(c0 + 3) * nctas + blockIdx.x] = t.w;
        }
        __syncthreads();
    }
}

// ---- small planes (N*H*W <= 16384 per channel: the 4 .. 16 px layers, and the (B, C) outputs of the style MLP): ONE pass.
// A channel belongs to one warp (<= 1024 elements) or one CTA; its threads walk the channel's N planes, write grad_in and
// fold their partial sums in a fixed order, so grad_bias needs neither the partial planes nor the fold launch that cost
// more than the pass itself at these sizes (0.15 of HBM at r = 16 in round 1).
template <typename T>
__global__ void __launch_bounds__(256) bias_act_bwd_small(T* __restrict__ gin, float* __restrict__ grad_bias,
                                                          const T* __restrict__ gout, const T* __restrict__ saved, int n,
                                                          int c, int hw, int tpc, float alpha, float scale) {
    __shared__ float s_part[8];
    const int per_cta = blockDim.x / tpc;
    const int ch = blockIdx.x * per_cta + threadIdx.x / tpc;
    const int t = threadIdx.x % tpc;
    const int total = n * hw;
    float acc = 0.f;
    if (ch < c) {
        for (int e = t; e < total; e += tpc) {
            const int i_n = e / hw, pos = e - i_n * hw;
            const long long idx = ((long long)i_n * c + ch) * hw + pos;
            const float g = Elem<T>::ld(gout + idx);
            const float r = (Elem<T>::ld(saved + idx) > 0.f ? g : g * alpha) * scale;
            Elem<T>::st(gin + idx, r);
            acc += r;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (tpc == 32) {
        if (t == 0 && ch < c) grad_bias[ch] = acc;
        return;
    }
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0 && ch < c) {
        float sum = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += s_part[w];
        grad_bias[ch] = sum;
    }
}

static inline int bwd_chunks(int64_t hw_vec) { return (int)ceil_div(hw_vec, 32 * kBwdUnroll); }

template <typename T>
static int launch_bias_act_bwd(void* gin, float* grad_bias, void* ws, const void* gout, const void* saved, int64_t n,
                               int64_t c, int64_t hw, float alpha, float scale, cudaStream_t s) {
    constexpr int N = Vec16<T>::N;
    const bool vec = (hw % N == 0) && aligned_to(gin, 16) && aligned_to(gout, 16) && aligned_to(saved, 16);
    const long long cap = (long long)kNumSMs * 16;
    const unsigned fold_blocks = (unsigned)ceil_div(c * 32, 256);
    if (n * hw <= 16384) {                       // small planes: one pass, no partials
        const int tpc = n * hw <= 1024 ? 32 : 256;
        const long long blocks = ceil_div(c, 256 / tpc);
        bias_act_bwd_small<T><<<(unsigned)blocks, 256, 0, s>>>((T*)gin, grad_bias, (const T*)gout, (const T*)saved, (int)n,
                                                                (int)c, (int)hw, tpc, alpha, scale);
        RICK_CHECK_LAUNCH();
        return RICK_OK;
    }
    if (vec) {
        const int64_t hw_vec = hw / N;
        const int chunks = bwd_chunks(hw_vec);
        const long long items = n * c * chunks;
        long long blocks = ceil_div(items, 8);
        if (blocks > cap) blocks = cap;
        bias_act_bwd_main<T><<<(unsigned)blocks, 256, 0, s>>>((T*)gin, (float*)ws, (const T*)gout, (const T*)saved,
                                                               n * c, hw_vec, chunks, alpha, scale);
        RICK_CHECK_LAUNCH();
        bias_grad_fold<float><<<fold_blocks, 256, 0, s>>>(grad_bias, (const float*)ws, (int)n, (int)c, chunks);
    } else {
        const long long total = n * c * hw;
        long long blocks = ceil_div(total, 256);
        if (blocks > cap) blocks = cap;
        bias_act_bwd_scalar<T><<<(unsigned)blocks, 256, 0, s>>>((T*)gin, (const T*)gout, (const T*)saved, total, alpha,
                                                                 scale);
        RICK_CHECK_LAUNCH();
        bias_grad_fold<T><<<fold_blocks, 256, 0, s>>>(grad_bias, (const T*)gin, (int)n, (int)c, hw);
    }
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

}  // namespace rick

extern "C" int rick_bias_act(void* out, const void* x, const void* bias, const void* ref, int64_t n, int64_t step_b,
                             int64_t size_b, int act, int grad, float alpha, float scale, int dtype,
                             rick_stream_t stream) {
    using namespace rick;
    if (!out || !x || n < 0) return RICK_ERR_INVALID_ARGUMENT;
    if (act != RICK_ACT_LINEAR && act != RICK_ACT_LRELU) return RICK_ERR_UNSUPPORTED;
    if (grad < 0 || grad > 2) return RICK_ERR_INVALID_ARGUMENT;
    if (bias && (step_b < 1 || size_b < 1 || size_b > 0x7fffffff)) return RICK_ERR_INVALID_ARGUMENT;
    if (dtype != RICK_F32 && dtype != RICK_BF16) return RICK_ERR_INVALID_ARGUMENT;
    if (n == 0) return RICK_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == RICK_F32)
        return launch_bias_act<float>(out, x, bias, ref, n, step_b, size_b, act, grad, alpha, scale, s);
    return launch_bias_act<__nv_bfloat16>(out, x, bias, ref, n, step_b, size_b, act, grad, alpha, scale, s);
}

extern "C" int64_t rick_bias_act_bwd_workspace(int64_t n, int64_t c, int64_t hw) {
    if (n < 1 || c < 1 || hw < 1) return 0;
    // sized for the widest case (fp32 vectors of 4): n*c*chunks floats
    return n * c * (int64_t)rick::bwd_chunks(rick::ceil_div(hw, 4)) * (int64_t)sizeof(float);
}

extern "C" int rick_bias_act_bwd(void* grad_in, float* grad_bias, void* workspace, const void* grad_out,
                                 const void* out_saved, int64_t n, int64_t c, int64_t hw, float alpha, float scale,
                                 int dtype, rick_stream_t stream) {
    using namespace rick;
    if (!grad_in || !grad_bias || !grad_out || !out_saved || !workspace) return RICK_ERR_INVALID_ARGUMENT;
    if (n < 1 || c < 1 || hw < 1 || c > 0x7fffffff || n > 0x7fffffff) return RICK_ERR_INVALID_ARGUMENT;
    if (dtype != RICK_F32 && dtype != RICK_BF16) return RICK_ERR_INVALID_ARGUMENT;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == RICK_F32)
        return launch_bias_act_bwd<float>(grad_in, grad_bias, workspace, grad_out, out_saved, n, c, hw, alpha, scale, s);
    return launch_bias_act_bwd<__nv_bfloat16>(grad_in, grad_bias, workspace, grad_out, out_saved, n, c, hw, alpha,
                                              scale, s);
}

extern "C" int64_t rick_bias_act_bwd_nhwc_workspace(int64_t rows, int64_t c) {
    if (rows < 1 || c < 1) return 0;
    return c * rick::ceil_div(rows, rick::kRowsPerCta) * (int64_t)sizeof(float);
}

extern "C" int rick_bias_act_bwd_nhwc(void* grad_in, float* grad_bias, void* workspace, const void* grad_out,
                                      const void* out_saved, int64_t rows, int64_t c, float alpha, float scale,
                                      rick_stream_t stream) {
    using namespace rick;
    if (!grad_in || !grad_bias || !grad_out || !out_saved || !workspace) return RICK_ERR_INVALID_ARGUMENT;
    if (rows < 1 || c < 4 || c > 0x7fffffff) return RICK_ERR_INVALID_ARGUMENT;
    if (c % 4 != 0) return RICK_ERR_UNSUPPORTED;
    if (!aligned_to(grad_in, 16) || !aligned_to(grad_out, 16) || !aligned_to(out_saved, 16)) return RICK_ERR_ALIGNMENT;
    const long long nctas = ceil_div(rows, kRowsPerCta);
    if (nctas > 0x7fffffffLL) return RICK_ERR_OVERFLOW;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int c4 = (int)(c / 4);
    bias_act_bwd_nhwc_main<<<(unsigned)nctas, 256, 256 * sizeof(float4), s>>>(
        (float*)grad_in, (float*)workspace, (const float*)grad_out, (const float*)out_saved, rows, c4, (int)nctas, alpha,
        scale);
    RICK_CHECK_LAUNCH();
    bias_grad_fold<float><<<(unsigned)ceil_div(c * 32, 256), 256, 0, s>>>(grad_bias, (const float*)workspace, 1, (int)c,
                                                                          nctas);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}
