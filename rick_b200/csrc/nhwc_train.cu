// Fused elementwise + reduction kernels around the differentiated (training-time) modulated convolution, channels-last
// fp32 (sm_100a).  They replace the chains of broadcast multiplies / adds and the ATen reductions that autograd builds
// for ModulatedConv2d + NoiseInjection + FusedLeakyReLU (model_probe_tune.py:246-251, 293-298, 342-348), which were ~30 %
// of the kernel time of an adaptation iteration in the round-1 profile:
//
//   rick_modulate_nhwc          y = x * s[b,c]
//   rick_modulate_bwd_nhwc      gx = gy * s[b,c]            gs[b,c]  = sum_hw gy * x
//   rick_styled_epilogue_nhwc   y = lrelu(a * d[b,c] + nw * noise[b,hw] + bias[c]) * scale
//   rick_styled_epilogue_bwd_nhwc  t = (y > 0 ? gy : gy*alpha) * scale
//                               ga = t * d[b,c]   gd[b,c] = sum_hw t*a   gbias[c] = sum_{b,hw} t   gnw = sum t*noise
//
// All are single-pass HBM-bound streams: 128-bit accesses along C, one CTA per 128 pixels of one sample, per-CTA
// partial sums folded in a fixed order (deterministic).
#include "common.cuh"

namespace rick {

namespace {

constexpr int kRows = 128;   // pixels per CTA

__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

// y = x * s   (grid-stride over float4 elements; rows = pixels per sample)
__global__ void __launch_bounds__(256) modulate_kernel(float4* __restrict__ y, const float4* __restrict__ x,
                                                       const float4* __restrict__ s, long long total4, int c4,
                                                       long long per_sample4) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += stride) {
        const int b = (int)(i / per_sample4);
        const int cg = (int)(i % c4);
        st_stream_f4(y + i, f4_mul(ld_stream_f4(x + i), __ldg(s + (long long)b * c4 + cg)));
    }
}

// MODE 0: modulate backward      in0 = gy, in1 = x            out = gy*s      p0[b,c] = sum gy*x
// MODE 1: styled-epilogue bwd    in0 = gy, in1 = y, in2 = a   out = t*d       p0 = sum t*a, p1 = sum t, pn = sum t*noise
template <int MODE>
__global__ void __launch_bounds__(256) colsum_bwd_kernel(float* __restrict__ out, float* __restrict__ p0,
                                                         float* __restrict__ p1, float* __restrict__ pn,
                                                         const float* __restrict__ in0, const float* __restrict__ in1,
                                                         const float* __restrict__ in2, const float* __restrict__ vec,
                                                         const float* __restrict__ noise, int hw, int c4, int nctas,
                                                         float alpha, float scale) {
    extern __shared__ float4 s_part[];            // [2][lanes][c4_tile]
    __shared__ float s_noise[8];
    const int b = blockIdx.y;
    const int c4_tile = c4 < 256 ? c4 : 256;
    const int lanes = 256 / c4_tile;
    const int cg_l = threadIdx.x % c4_tile, rl = threadIdx.x / c4_tile;
    const int row0 = blockIdx.x * kRows;
    const int row_end = min(row0 + kRows, hw);
    float nacc = 0.f;
    for (int cg0 = 0; cg0 < c4; cg0 += c4_tile) {
        const int cg = cg0 + cg_l;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        if (rl < lanes && cg < c4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(vec) + (long long)b * c4 + cg);   // s or d
            for (int r = row0 + rl; r < row_end; r += lanes) {
                const long long i = ((long long)b * hw + r) * c4 + cg;
                const float4 g = ld_stream_f4(reinterpret_cast<const float4*>(in0) + i);
                const float4 u = ld_stream_f4(reinterpret_cast<const float4*>(in1) + i);
                if (MODE == 0) {
                    st_stream_f4(reinterpret_cast<float4*>(out) + i, f4_mul(g, v));
                    a0.x += g.x * u.x, a0.y += g.y * u.y, a0.z += g.z * u.z, a0.w += g.w * u.w;
                } else {
                    const float4 a = ld_stream_f4(reinterpret_cast<const float4*>(in2) + i);
                    float4 t;
                    t.x = (u.x > 0.f ? g.x : g.x * alpha) * scale, t.y = (u.y > 0.f ? g.y : g.y * alpha) * scale;
                    t.z = (u.z > 0.f ? g.z : g.z * alpha) * scale, t.w = (u.w > 0.f ? g.w : g.w * alpha) * scale;
                    st_stream_f4(reinterpret_cast<float4*>(out) + i, f4_mul(t, v));
                    a0.x += t.x * a.x, a0.y += t.y * a.y, a0.z += t.z * a.z, a0.w += t.w * a.w;
                    a1.x += t.x, a1.y += t.y, a1.z += t.z, a1.w += t.w;
                    if (noise) nacc += (t.x + t.y + t.z + t.w) * __ldg(noise + (long long)b * hw + r);
                }
            }
        }
        if (rl < lanes) {
            s_part[rl * c4_tile + cg_l] = a0;
            if (MODE == 1) s_part[(lanes + rl) * c4_tile + cg_l] = a1;
        }
        __syncthreads();
        if (rl == 0 && cg < c4) {
            float4 t0 = s_part[cg_l], t1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (MODE == 1) t1 = s_part[lanes * c4_tile + cg_l];
            for (int l = 1; l < lanes; ++l) {
                const float4 u0 = s_part[l * c4_tile + cg_l];
                t0.x += u0.x, t0.y += u0.y, t0.z += u0.z, t0.w += u0.w;
                if (MODE == 1) {
                    const float4 u1 = s_part[(lanes + l) * c4_tile + cg_l];
                    t1.x += u1.x, t1.y += u1.y, t1.z += u1.z, t1.w += u1.w;
                }
            }
            const long long base = ((long long)b * c4 * 4 + (long long)cg * 4) * nctas + blockIdx.x;   // [(b*C + c)][cta]
            p0[base] = t0.x, p0[base + nctas] = t0.y, p0[base + 2LL * nctas] = t0.z, p0[base + 3LL * nctas] = t0.w;
            if (MODE == 1) {
                p1[base] = t1.x, p1[base + nctas] = t1.y, p1[base + 2LL * nctas] = t1.z, p1[base + 3LL * nctas] = t1.w;
            }
        }
        __syncthreads();
    }
    if (MODE == 1 && pn) {            // block-wide sum of the noise-weight partial, fixed order
        nacc = warp_sum(nacc);
        if ((threadIdx.x & 31) == 0) s_noise[threadIdx.x >> 5] = nacc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int wv = 0; wv < 8; ++wv) t += s_noise[wv];
            pn[(long long)b * nctas + blockIdx.x] = t;
        }
    }
}

// y = lrelu(a*d + nw*noise + bias) * scale
__global__ void __launch_bounds__(256) styled_epilogue_kernel(float4* __restrict__ y, const float4* __restrict__ a,
                                                              const float4* __restrict__ d,
                                                              const float* __restrict__ noise,
                                                              const float* __restrict__ noise_w,
                                                              const float4* __restrict__ bias, long long total4, int c4,
                                                              long long per_sample4, float alpha, float scale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float nw = noise ? __ldg(noise_w) : 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += stride) {
        const int b = (int)(i / per_sample4);
        const int cg = (int)(i % c4);
        const float4 v = ld_stream_f4(a + i);
        const float4 dm = d ? __ldg(d + (long long)b * c4 + cg) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 bs = bias ? __ldg(bias + cg) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float nz = noise ? nw * __ldg(noise + i / c4) : 0.f;     // pixel index = i / c4 (b-major)
        float4 r;
        r.x = fmaf(v.x, dm.x, nz) + bs.x, r.y = fmaf(v.y, dm.y, nz) + bs.y;
        r.z = fmaf(v.z, dm.z, nz) + bs.z, r.w = fmaf(v.w, dm.w, nz) + bs.w;
        r.x = (r.x > 0.f ? r.x : r.x * alpha) * scale, r.y = (r.y > 0.f ? r.y : r.y * alpha) * scale;
        r.z = (r.z > 0.f ? r.z : r.z * alpha) * scale, r.w = (r.w > 0.f ? r.w : r.w * alpha) * scale;
        st_stream_f4(y + i, r);
    }
}

static bool ok16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace rick

extern "C" int rick_modulate_nhwc(void* y, const void* x, const float* s, int batch, int64_t hw, int channels,
                                  rick_stream_t stream) {
    using namespace rick;
    if (!y || !x || !s || batch < 1 || hw < 1 || channels < 4) return RICK_ERR_INVALID_ARGUMENT;
    if (channels % 4) return RICK_ERR_UNSUPPORTED;
    if (!ok16(y) || !ok16(x) || !ok16(s)) return RICK_ERR_ALIGNMENT;
    const int c4 = channels / 4;
    const long long per4 = hw * c4, total4 = per4 * batch;
    long long blocks = ceil_div(total4, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    modulate_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        (float4*)y, (const float4*)x, (const float4*)s, total4, c4, per4);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

extern "C" int64_t rick_colsum_workspace(int batch, int64_t hw, int channels) {
    if (batch < 1 || hw < 1 || channels < 1) return 0;
    const int64_t nctas = rick::ceil_div(hw, rick::kRows);
    return ((int64_t)2 * batch * channels * nctas + (int64_t)batch * nctas) * (int64_t)sizeof(float);
}

extern "C" int rick_modulate_bwd_nhwc(void* gx, float* gs, void* workspace, const void* gy, const void* x, const float* s,
                                      int batch, int64_t hw, int channels, rick_stream_t stream) {
    using namespace rick;
    if (!gx || !gs || !workspace || !gy || !x || !s || batch < 1 || hw < 1 || channels < 4) return RICK_ERR_INVALID_ARGUMENT;
    if (channels % 4) return RICK_ERR_UNSUPPORTED;
    if (hw > 0x7fffffffLL || batch > 65535) return RICK_ERR_OVERFLOW;
    if (!ok16(gx) || !ok16(gy) || !ok16(x) || !ok16(s)) return RICK_ERR_ALIGNMENT;
    const int c4 = channels / 4;
    const int nctas = (int)ceil_div(hw, kRows);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* p0 = static_cast<float*>(workspace);
    colsum_bwd_kernel<0><<<dim3(nctas, batch), 256, 2 * 256 * sizeof(float4), st>>>(
        (float*)gx, p0, nullptr, nullptr, (const float*)gy, (const float*)x, nullptr, s, nullptr, (int)hw, c4, nctas, 0.f,
        1.f);
    RICK_CHECK_LAUNCH();
    bias_grad_fold<float><<<(unsigned)ceil_div((long long)batch * channels * 32, 256), 256, 0, st>>>(
        gs, p0, 1, batch * channels, nctas);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

extern "C" int rick_styled_epilogue_nhwc(void* y, const void* a, const float* demod, const float* noise,
                                         const float* noise_weight, const float* bias, int batch, int64_t hw,
                                         int channels, float alpha, float scale, rick_stream_t stream) {
    using namespace rick;
    if (!y || !a || batch < 1 || hw < 1 || channels < 4) return RICK_ERR_INVALID_ARGUMENT;
    if (channels % 4) return RICK_ERR_UNSUPPORTED;
    if (noise && !noise_weight) return RICK_ERR_INVALID_ARGUMENT;
    if (!ok16(y) || !ok16(a) || (demod && !ok16(demod)) || (bias && !ok16(bias))) return RICK_ERR_ALIGNMENT;
    const int c4 = channels / 4;
    const long long per4 = hw * c4, total4 = per4 * batch;
    long long blocks = ceil_div(total4, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    styled_epilogue_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        (float4*)y, (const float4*)a, (const float4*)demod, noise, noise_weight, (const float4*)bias, total4, c4, per4,
        alpha, scale);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

extern "C" int rick_styled_epilogue_bwd_nhwc(void* ga, float* gdemod, float* gbias, float* gnoise_weight, void* workspace,
                                             const void* gy, const void* y, const void* a, const float* demod,
                                             const float* noise, int batch, int64_t hw, int channels, float alpha,
                                             float scale, rick_stream_t stream) {
    using namespace rick;
    if (!ga || !gdemod || !gbias || !workspace || !gy || !y || !a || !demod || batch < 1 || hw < 1 || channels < 4)
        return RICK_ERR_INVALID_ARGUMENT;
    if (channels % 4) return RICK_ERR_UNSUPPORTED;
    if (noise && !gnoise_weight) return RICK_ERR_INVALID_ARGUMENT;
    if (hw > 0x7fffffffLL || batch > 65535) return RICK_ERR_OVERFLOW;
    if (!ok16(ga) || !ok16(gy) || !ok16(y) || !ok16(a) || !ok16(demod)) return RICK_ERR_ALIGNMENT;
    const int c4 = channels / 4;
    const int nctas = (int)ceil_div(hw, kRows);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* p0 = static_cast<float*>(workspace);
    float* p1 = p0 + (size_t)batch * channels * nctas;
    float* pn = p1 + (size_t)batch * channels * nctas;
    colsum_bwd_kernel<1><<<dim3(nctas, batch), 256, 2 * 256 * sizeof(float4), st>>>(
        (float*)ga, p0, p1, noise ? pn : nullptr, (const float*)gy, (const float*)y, (const float*)a, demod, noise, (int)hw,
        c4, nctas, alpha, scale);
    RICK_CHECK_LAUNCH();
    bias_grad_fold<float><<<(unsigned)ceil_div((long long)batch * channels * 32, 256), 256, 0, st>>>(
        gdemod, p0, 1, batch * channels, nctas);
    RICK_CHECK_LAUNCH();
    bias_grad_fold<float><<<(unsigned)ceil_div((long long)channels * 32, 256), 256, 0, st>>>(gbias, p1, batch, channels,
                                                                                               nctas);
    RICK_CHECK_LAUNCH();
    if (noise) {
        bias_grad_fold<float><<<1, 256, 0, st>>>(gnoise_weight, pn, 1, 1, (long long)batch * nctas);
        RICK_CHECK_LAUNCH();
    }
    return RICK_OK;
}
