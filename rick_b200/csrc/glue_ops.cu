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

// ------------------------------------------------------------------------------------------------------------------
// The discriminator's from-RGB layer: ConvLayer(3, C, 1) = EqualConv2d 1x1 (no bias) + FusedLeakyReLU
// (model_probe_tune.py:595-641, 676).  With K = 3 input channels it is an outer product per pixel, not a GEMM: as
// library / element-wise calls it cost 3 broadcast multiplies, 2 adds, a layout copy and the bias-act pass over the
// (B, C, 256, 256) output (0.55 ms at batch 4), and its data gradient three more broadcast multiply + reduce pairs.
//
//   from_rgb_fwd      y[b,p,co] = lrelu( w_scale * sum_c img[b,c,p] * w[co,c] + bias[co] ) * act_scale       (NCHW -> NHWC)
//   from_rgb_bwd_data gimg[b,c,p] = w_scale * sum_co w[co,c] * t[b,p,co],  t = (y > 0 ? g : alpha*g) * act_scale
//
// Both are one pass over the feature map (write-bound / read-bound).
namespace rick {
namespace {

__global__ void __launch_bounds__(256)
from_rgb_fwd_kernel(float* __restrict__ y, const float* __restrict__ img, const float* __restrict__ w,
                    const float* __restrict__ bias, long long pixels_per_sample, int batch, int cin, int cout, float w_scale,
                    int act, float alpha, float act_scale) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = pixels_per_sample * batch;
    for (int cq = lane * 4; cq < cout; cq += 128) {           // this lane's 4 output channels (cout % 4 == 0)
        float wr[4][4];
        float br[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            br[j] = bias ? __ldg(bias + cq + j) : 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) wr[j][c] = c < cin ? __ldg(w + (long long)(cq + j) * cin + c) * w_scale : 0.f;
        }
        for (long long p = warp; p < total; p += n_warps) {
            const long long b = p / pixels_per_sample, q = p - b * pixels_per_sample;
            const float* src = img + b * cin * pixels_per_sample + q;
            float x[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) x[c] = c < cin ? __ldg(src + c * pixels_per_sample) : 0.f;
            float4 o;
            float* op = reinterpret_cast<float*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = fmaf(x[0], wr[j][0], fmaf(x[1], wr[j][1], fmaf(x[2], wr[j][2], fmaf(x[3], wr[j][3], br[j]))));
                if (act) v = (v > 0.f ? v : v * alpha) * act_scale;
                op[j] = v;
            }
            st_stream_f4(reinterpret_cast<float4*>(y + p * cout + cq), o);
        }
    }
}

__global__ void __launch_bounds__(256)
from_rgb_bwd_data_kernel(float* __restrict__ gimg, const float* __restrict__ g, const float* __restrict__ y,
                         const float* __restrict__ w, long long pixels_per_sample, int batch, int cin, int cout,
                         float w_scale, int act, float alpha, float act_scale) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long total = pixels_per_sample * batch;
    for (long long p = warp; p < total; p += n_warps) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int cq = lane * 4; cq < cout; cq += 128) {
            const float4 gv = ld_stream_f4(reinterpret_cast<const float4*>(g + p * cout + cq));
            float t[4] = {gv.x, gv.y, gv.z, gv.w};
            if (act) {
                const float4 yv = ld_stream_f4(reinterpret_cast<const float4*>(y + p * cout + cq));
                const float ys[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) t[j] = (ys[j] > 0.f ? t[j] : t[j] * alpha) * act_scale;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < cin) acc[c] = fmaf(t[j], __ldg(w + (long long)(cq + j) * cin + c), acc[c]);
        }
        const long long b = p / pixels_per_sample, q = p - b * pixels_per_sample;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < cin) {
                const float s = warp_sum(acc[c]);
                if (lane == 0) gimg[(b * cin + c) * pixels_per_sample + q] = s * w_scale;
            }
        }
    }
}

}  // namespace
}  // namespace rick

extern "C" int rick_from_rgb_fwd(float* y, const float* img, const float* w, const float* bias, int batch, int64_t pixels,
                                 int cin, int cout, float w_scale, int act, float alpha, float act_scale,
                                 rick_stream_t stream) {
    using namespace rick;
    if (!y || !img || !w || batch < 1 || pixels < 1) return RICK_ERR_INVALID_ARGUMENT;
    if (cin < 1 || cin > 4 || cout < 4 || cout % 4 != 0) return RICK_ERR_UNSUPPORTED;
    if (!aligned_to(y, 16)) return RICK_ERR_ALIGNMENT;
    long long blocks = ceil_div(pixels * batch, 8 * 4);       // a warp handles ~4 pixels
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    from_rgb_fwd_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, img, w, bias, pixels, batch, cin,
                                                                                          cout, w_scale, act, alpha, act_scale);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

extern "C" int rick_from_rgb_bwd_data(float* gimg, const float* g, const float* y, const float* w, int batch, int64_t pixels,
                                      int cin, int cout, float w_scale, int act, float alpha, float act_scale,
                                      rick_stream_t stream) {
    using namespace rick;
    if (!gimg || !g || !w || (act && !y) || batch < 1 || pixels < 1) return RICK_ERR_INVALID_ARGUMENT;
    if (cin < 1 || cin > 4 || cout < 4 || cout % 4 != 0) return RICK_ERR_UNSUPPORTED;
    if (!aligned_to(g, 16) || (y && !aligned_to(y, 16))) return RICK_ERR_ALIGNMENT;
    long long blocks = ceil_div(pixels * batch, 8 * 4);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    from_rgb_bwd_data_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(gimg, g, y, w, pixels, batch, cin,
                                                                                               cout, w_scale, act, alpha,
                                                                                               act_scale);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

// -------------------------------------------------------------------------------------------------------------------
// demodulation table of a ModulatedConv2d (model_probe_tune.py:246-251 in its algebraic form):
//     demod[b][co] = rsqrt(scale2 * sum_ci s[b][ci]^2 * wsq[co][ci] + eps),     s_out[b][ci] = s[b][ci] * s_scale
// As module code this is pow / mm / mul / add / rsqrt / mul on (B, 512) tensors: six launches per layer and pass, the mm a
// 10 us cuBLAS GEMV (0.65 ms of a 13 ms iteration in the round-2 glue profile).  One warp per output channel here.
// -------------------------------------------------------------------------------------------------------------------
namespace rick {
namespace {
constexpr int kDemodMaxBatch = 8;

__global__ void __launch_bounds__(256)
demod_fwd_kernel(float* __restrict__ demod, float* __restrict__ s_out, const float* __restrict__ s,
                 const float* __restrict__ wsq, int batch, int cin, int cout, float scale2, float eps, float s_scale) {
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s_out) {                                   // the scaled style: every thread of the grid takes a few elements
        for (int i = blockIdx.x * 256 + threadIdx.x; i < batch * cin; i += gridDim.x * 256) s_out[i] = s[i] * s_scale;
    }
    if (warp >= cout) return;
    float acc[kDemodMaxBatch];
#pragma unroll
    for (int b = 0; b < kDemodMaxBatch; ++b) acc[b] = 0.f;
    const float* wrow = wsq + (size_t)warp * cin;
    for (int ci = lane; ci < cin; ci += 32) {
        const float w = __ldg(wrow + ci);
#pragma unroll
        for (int b = 0; b < kDemodMaxBatch; ++b)
            if (b < batch) {
                const float v = __ldg(s + (size_t)b * cin + ci);
                acc[b] = fmaf(v * v, w, acc[b]);
            }
    }
#pragma unroll
    for (int b = 0; b < kDemodMaxBatch; ++b) {
        if (b < batch) {
            float a = acc[b];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) demod[(size_t)b * cout + warp] = rsqrtf(fmaf(a, scale2, eps));
        }
    }
}

// q[b][co] = -0.5 * demod^3 * scale2 * g_demod;   g_wsq[co][ci] = sum_b q[b][co] * s[b][ci]^2
__global__ void __launch_bounds__(256)
demod_bwd_wsq_kernel(float* __restrict__ g_wsq, const float* __restrict__ g_demod, const float* __restrict__ demod,
                     const float* __restrict__ s, int batch, int cin, int cout, float scale2) {
    const long long n = (long long)cout * cin;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        const int co = (int)(i / cin), ci = (int)(i - (long long)co * cin);
        float a = 0.f;
        for (int b = 0; b < batch; ++b) {
            const float d = __ldg(demod + (size_t)b * cout + co);
            const float q = -0.5f * d * d * d * scale2 * __ldg(g_demod + (size_t)b * cout + co);
            const float v = __ldg(s + (size_t)b * cin + ci);
            a = fmaf(q, v * v, a);
        }
        g_wsq[i] = a;
    }
}

// g_s[b][ci] = 2 * s[b][ci] * sum_co q[b][co] * wsq[co][ci] + s_scale * g_sout[b][ci]
// CTA = 32 consecutive ci x 8 warps striding over co.  q is staged in shared memory once; every warp keeps 8 rows of wsq
// in flight (the first version walked co one dependent load at a time: 36 us per layer, more than the convolution it
// serves at 4 .. 16 px).  Warp partials are folded through shared memory in a fixed order.
__global__ void __launch_bounds__(256)
demod_bwd_s_kernel(float* __restrict__ g_s, const float* __restrict__ g_demod, const float* __restrict__ g_sout,
                   const float* __restrict__ demod, const float* __restrict__ s, const float* __restrict__ wsq, int batch,
                   int cin, int cout, float scale2, float s_scale) {
    extern __shared__ float s_q[];                         // [batch][cout]
    __shared__ float part[8][kDemodMaxBatch][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < batch * cout; i += 256) {
        const float d = __ldg(demod + i);
        s_q[i] = -0.5f * d * d * d * scale2 * __ldg(g_demod + i);
    }
    __syncthreads();
    const int ci = blockIdx.x * 32 + lane;
    float acc[kDemodMaxBatch];
#pragma unroll
    for (int b = 0; b < kDemodMaxBatch; ++b) acc[b] = 0.f;
    for (int co0 = warp * 8; co0 < cout; co0 += 64) {
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = (ci < cin && co0 + j < cout) ? __ldg(wsq + (size_t)(co0 + j) * cin + ci) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (co0 + j < cout) {
#pragma unroll
                for (int b = 0; b < kDemodMaxBatch; ++b)
                    if (b < batch) acc[b] = fmaf(s_q[b * cout + co0 + j], w[j], acc[b]);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < kDemodMaxBatch; ++b) part[warp][b][lane] = acc[b];
    __syncthreads();
    if (warp < batch && ci < cin) {                // warp b finishes sample b
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += part[w][warp][lane];
        const size_t idx = (size_t)warp * cin + ci;
        float r = 2.f * __ldg(s + idx) * a;
        if (g_sout) r = fmaf(s_scale, __ldg(g_sout + idx), r);
        g_s[idx] = r;
    }
}

__global__ void __launch_bounds__(256)
add_scale_kernel(float4* __restrict__ out, const float4* __restrict__ a, const float4* __restrict__ b, float scale,
                 long long n4) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) {
        const float4 x = ld_stream_f4(a + i), y = ld_stream_f4(b + i);
        st_stream_f4(out + i, make_float4((x.x + y.x) * scale, (x.y + y.y) * scale, (x.z + y.z) * scale, (x.w + y.w) * scale));
    }
}
}  // namespace
}  // namespace rick

extern "C" int rick_demod_fwd(float* demod, float* s_out, const float* s, const float* wsq, int batch, int cin, int cout,
                              float scale2, float eps, float s_scale, rick_stream_t stream) {
    using namespace rick;
    if (!demod || !s || !wsq || batch < 1 || cin < 1 || cout < 1) return RICK_ERR_INVALID_ARGUMENT;
    if (batch > kDemodMaxBatch) return RICK_ERR_UNSUPPORTED;
    const unsigned blocks = (unsigned)ceil_div((long long)cout * 32, 256);
    demod_fwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(demod, s_out, s, wsq, batch, cin, cout, scale2,
                                                                          eps, s_scale);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

extern "C" int rick_demod_bwd(float* g_s, float* g_wsq, const float* g_demod, const float* g_sout, const float* demod,
                              const float* s, const float* wsq, int batch, int cin, int cout, float scale2, float s_scale,
                              rick_stream_t stream) {
    using namespace rick;
    if (!g_demod || !demod || !s || !wsq || batch < 1 || cin < 1 || cout < 1) return RICK_ERR_INVALID_ARGUMENT;
    if (batch > kDemodMaxBatch) return RICK_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (g_wsq) {
        long long blocks = ceil_div((long long)cout * cin, 256);
        if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        demod_bwd_wsq_kernel<<<(unsigned)blocks, 256, 0, st>>>(g_wsq, g_demod, demod, s, batch, cin, cout, scale2);
        RICK_CHECK_LAUNCH();
    }
    if (g_s) {
        if ((size_t)batch * cout * sizeof(float) > 40 * 1024) return RICK_ERR_UNSUPPORTED;
        demod_bwd_s_kernel<<<(unsigned)ceil_div(cin, 32), 256, (size_t)batch * cout * sizeof(float), st>>>(
            g_s, g_demod, g_sout, demod, s, wsq, batch, cin, cout, scale2, s_scale);
        RICK_CHECK_LAUNCH();
    }
    return RICK_OK;
}

extern "C" int rick_add_scale(float* out, const float* a, const float* b, float scale, int64_t n, rick_stream_t stream) {
    using namespace rick;
    if (!out || !a || !b || n < 0) return RICK_ERR_INVALID_ARGUMENT;
    if (n % 4 != 0) return RICK_ERR_UNSUPPORTED;
    if (!aligned_to(out, 16) || !aligned_to(a, 16) || !aligned_to(b, 16)) return RICK_ERR_ALIGNMENT;
    if (n == 0) return RICK_OK;
    long long blocks = ceil_div(n / 4, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    add_scale_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<float4*>(out), reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), scale, n / 4);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}
