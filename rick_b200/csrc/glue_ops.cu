// Small fused kernels that stand where the module layer would otherwise run chains of ATen element-wise / reduction
// launches around the convolutions (sm_100a).  Each is a single memory-bound pass.
//
//   rick_weight_sqsum    wsq[co][ci] = sum_taps W[co][ci][tap]^2      the per-layer table ModulatedConv2d's demodulation
//                        is computed from (model_probe_tune.py:249-251 in its algebraic form:
//                        demod[b,co] = rsqrt(sum_ci s[b,ci]^2 * wsq[co,ci] + eps)), for any weight strides
#include "common.cuh"

namespace rick {
namespace {

__global__ void __launch_bounds__(256)
weight_sqsum_kernel(float* __restrict__ out, const float* __restrict__ w, int cout, int cin, int taps, long long s_co,
                    long long s_ci, long long s_tap) {
    const long long n = (long long)cout * cin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % cin);
        const int co = (int)(i / cin);
        const float* p = w + co * s_co + ci * s_ci;
        float acc = 0.f;
        for (int t = 0; t < taps; ++t) {
            const float v = __ldg(p + t * s_tap);
            acc = fmaf(v, v, acc);
        }
        out[i] = acc;
    }
}

}  // namespace
}  // namespace rick

extern "C" int rick_weight_sqsum(float* out, const float* w, int cout, int cin, int taps, int64_t stride_co,
                                 int64_t stride_ci, int64_t stride_tap, rick_stream_t stream) {
    using namespace rick;
    if (!out || !w || cout < 1 || cin < 1 || taps < 1) return RICK_ERR_INVALID_ARGUMENT;
    long long blocks = ceil_div((long long)cout * cin, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    weight_sqsum_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, w, cout, cin, taps, stride_co,
                                                                                          stride_ci, stride_tap);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Small-batch linear layers, many at once.
//
// The generator evaluates 8 mapping-network layers and one style -> channel modulation layer per ModulatedConv2d (20 at
// 256 px) on a batch of 2-4 rows: as library calls that is one GEMV-shaped cuBLAS launch plus a bias / scale / activation
// launch or two per layer (98 + ~100 launches, ~1 ms per adaptation iteration in the round-2 launch list).  Here one
// launch serves a whole table of layers (EqualLinear semantics, model_probe_tune.py:139-173):
//
//     y_l[b, r] = act( w_scale_l * sum_k W_l[r, k] * x_l[b, k]  +  b_scale_l * bias_l[r] )
//
// A warp owns one output row r: it streams W_l[r, :] once (the only real traffic: 2 KB per row) and keeps one accumulator
// per batch row.  `pixelnorm` normalises x first (PixelNorm, model_probe_tune.py:21-26).
namespace rick {
namespace {

constexpr int kMaxLinear = 32;        // layers per launch
constexpr int kMaxLinBatch = 8;       // batch rows

struct LinearTable {
    const float* w[kMaxLinear];
    const float* bias[kMaxLinear];
    const float* x[kMaxLinear];       // (batch, in_dim) rows x_stride[l] elements apart
    float* y[kMaxLinear];             // (batch, out_dim[l]) contiguous
    long long x_stride[kMaxLinear];
    int out_dim[kMaxLinear];
    int row_end[kMaxLinear];          // prefix sum of out_dim: warp -> layer lookup
    float w_scale[kMaxLinear], b_scale[kMaxLinear];
    int count, batch, in_dim, act, pixelnorm;
    float alpha, act_scale;
};

__device__ __forceinline__ int find_layer(const int* row_end, int count, int row) {
    int l = 0;
    while (l + 1 < count && row >= row_end[l]) ++l;
    return l;
}

__global__ void __launch_bounds__(256) linear_multi_kernel(const __grid_constant__ LinearTable t) {
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    if (row >= t.row_end[t.count - 1]) return;
    const int l = find_layer(t.row_end, t.count, row);
    const int r = row - (l ? t.row_end[l - 1] : 0);
    const float4* __restrict__ wrow = reinterpret_cast<const float4*>(t.w[l] + (long long)r * t.in_dim);
    float acc[kMaxLinBatch], nrm[kMaxLinBatch];
#pragma unroll
    for (int b = 0; b < kMaxLinBatch; ++b) acc[b] = 0.f, nrm[b] = 0.f;
    const int n4 = t.in_dim >> 2;
    for (int j = lane; j < n4; j += 32) {
        const float4 wv = ld_stream_f4(wrow + j);
#pragma unroll
        for (int b = 0; b < kMaxLinBatch; ++b) {
            if (b < t.batch) {
                const float4 xv = __ldg(reinterpret_cast<const float4*>(t.x[l] + b * t.x_stride[l]) + j);
                acc[b] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[b]))));
                if (t.pixelnorm) nrm[b] = fmaf(xv.x, xv.x, fmaf(xv.y, xv.y, fmaf(xv.z, xv.z, fmaf(xv.w, xv.w, nrm[b]))));
            }
        }
    }
    const float bias = t.bias[l] ? __ldg(t.bias[l] + r) * t.b_scale[l] : 0.f;
#pragma unroll
    for (int b = 0; b < kMaxLinBatch; ++b) {
        if (b < t.batch) {
            float v = warp_sum(acc[b]);
            if (t.pixelnorm) v *= rsqrtf(warp_sum(nrm[b]) / (float)t.in_dim + 1e-8f);
            v = fmaf(v, t.w_scale[l], bias);
            if (t.act) v = (v > 0.f ? v : v * t.alpha) * t.act_scale;
            if (lane == 0) t.y[l][(long long)b * t.out_dim[l] + r] = v;
        }
    }
}

// gW_l[r, k] = w_scale_l * sum_b gy_l[b, r] * x_l[b, k];   gbias_l[r] = b_scale_l * sum_b gy_l[b, r]
struct LinearGradTable {
    float* gw[kMaxLinear];
    float* gbias[kMaxLinear];
    const float* gy[kMaxLinear];
    const float* x[kMaxLinear];
    long long x_stride[kMaxLinear];
    int out_dim[kMaxLinear];
    int row_end[kMaxLinear];
    float w_scale[kMaxLinear], b_scale[kMaxLinear];
    int count, batch, in_dim;
};

__global__ void __launch_bounds__(256) linear_multi_wgrad_kernel(const __grid_constant__ LinearGradTable t) {
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    if (row >= t.row_end[t.count - 1]) return;
    const int l = find_layer(t.row_end, t.count, row);
    const int r = row - (l ? t.row_end[l - 1] : 0);
    float g[kMaxLinBatch];
    float gsum = 0.f;
#pragma unroll
    for (int b = 0; b < kMaxLinBatch; ++b) {
        g[b] = b < t.batch ? __ldg(t.gy[l] + (long long)b * t.out_dim[l] + r) : 0.f;
        gsum += g[b];
    }
    if (t.gbias[l] && lane == 0) t.gbias[l][r] = gsum * t.b_scale[l];
    if (!t.gw[l]) return;
    float4* __restrict__ dst = reinterpret_cast<float4*>(t.gw[l] + (long long)r * t.in_dim);
    const int n4 = t.in_dim >> 2;
    const float ws = t.w_scale[l];
    for (int j = lane; j < n4; j += 32) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int b = 0; b < kMaxLinBatch; ++b) {
            if (b < t.batch) {
                const float4 xv = __ldg(reinterpret_cast<const float4*>(t.x[l] + b * t.x_stride[l]) + j);
                a.x = fmaf(g[b], xv.x, a.x), a.y = fmaf(g[b], xv.y, a.y), a.z = fmaf(g[b], xv.z, a.z), a.w = fmaf(g[b], xv.w, a.w);
            }
        }
        a.x *= ws, a.y *= ws, a.z *= ws, a.w *= ws;
        st_stream_f4(dst + j, a);
    }
}

}  // namespace
}  // namespace rick

extern "C" int rick_linear_multi(float* const* y, const float* const* w, const float* const* bias, const float* const* x,
                                 const int64_t* x_stride, const int* out_dim, const float* w_scale, const float* b_scale,
                                 int count, int batch, int in_dim, int act, float alpha, float act_scale, int pixelnorm,
                                 rick_stream_t stream) {
    using namespace rick;
    if (!y || !w || !bias || !x || !x_stride || !out_dim || !w_scale || !b_scale) return RICK_ERR_INVALID_ARGUMENT;
    if (count < 0 || batch < 1 || batch > kMaxLinBatch || in_dim < 4 || in_dim % 4 != 0) return RICK_ERR_UNSUPPORTED;
    for (int base = 0; base < count; base += kMaxLinear) {
        LinearTable t{};
        const int m = count - base < kMaxLinear ? count - base : kMaxLinear;
        int rows = 0;
        for (int i = 0; i < m; ++i) {
            const int k = base + i;
            if (!y[k] || !w[k] || !x[k] || out_dim[k] < 1 || x_stride[k] % 4 != 0) return RICK_ERR_INVALID_ARGUMENT;
            if (!aligned_to(w[k], 16) || !aligned_to(x[k], 16)) return RICK_ERR_ALIGNMENT;
            t.w[i] = w[k], t.bias[i] = bias[k], t.x[i] = x[k], t.y[i] = y[k], t.x_stride[i] = x_stride[k];
            t.out_dim[i] = out_dim[k], t.w_scale[i] = w_scale[k], t.b_scale[i] = b_scale[k];
            rows += out_dim[k];
            t.row_end[i] = rows;
        }
        t.count = m, t.batch = batch, t.in_dim = in_dim, t.act = act, t.pixelnorm = pixelnorm, t.alpha = alpha;
        t.act_scale = act_scale;
        if (!m) continue;
        linear_multi_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(t);
        RICK_CHECK_LAUNCH();
    }
    return RICK_OK;
}

extern "C" int rick_linear_multi_wgrad(float* const* gw, float* const* gbias, const float* const* gy, const float* const* x,
                                       const int64_t* x_stride, const int* out_dim, const float* w_scale,
                                       const float* b_scale, int count, int batch, int in_dim, rick_stream_t stream) {
    using namespace rick;
    if (!gw || !gbias || !gy || !x || !x_stride || !out_dim || !w_scale || !b_scale) return RICK_ERR_INVALID_ARGUMENT;
    if (count < 0 || batch < 1 || batch > kMaxLinBatch || in_dim < 4 || in_dim % 4 != 0) return RICK_ERR_UNSUPPORTED;
    for (int base = 0; base < count; base += kMaxLinear) {
        LinearGradTable t{};
        const int m = count - base < kMaxLinear ? count - base : kMaxLinear;
        int rows = 0;
        for (int i = 0; i < m; ++i) {
            const int k = base + i;
            if (!gy[k] || !x[k] || out_dim[k] < 1 || x_stride[k] % 4 != 0) return RICK_ERR_INVALID_ARGUMENT;
            if ((gw[k] && !aligned_to(gw[k], 16)) || !aligned_to(x[k], 16)) return RICK_ERR_ALIGNMENT;
            t.gw[i] = gw[k], t.gbias[i] = gbias[k], t.gy[i] = gy[k], t.x[i] = x[k], t.x_stride[i] = x_stride[k];
            t.out_dim[i] = out_dim[k], t.w_scale[i] = w_scale[k], t.b_scale[i] = b_scale[k];
            rows += out_dim[k];
            t.row_end[i] = rows;
        }
        t.count = m, t.batch = batch, t.in_dim = in_dim;
        if (!m) continue;
        linear_multi_wgrad_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(t);
        RICK_CHECK_LAUNCH();
    }
    return RICK_OK;
}
