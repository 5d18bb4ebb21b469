// Channels-last (NHWC) companions of the tcgen05 convolution for the generator's fused forward path (sm_100a):
//
//   rick_blur_nhwc    4x4 FIR (up = down = 1, any pad) over (B, H, W, C) with the StyledConv epilogue fused in:
//                     out = lrelu( blur(x) * demod[b,c] + noise_w * noise[b,y,x] + bias[c] ) * scale, and optionally
//                     out2 = out * s_next[b,c] (the next layer's pre-modulated input).  This is the Blur after the
//                     stride-2 transposed convolution (model_probe_tune.py:268) + NoiseInjection (:293-298) +
//                     FusedLeakyReLU in ONE pass over the activation instead of four.
//   rick_to_rgb_nhwc  ToRGB (model_probe_tune.py:361-370): 1x1 modulated conv to 3 channels (no demodulation) + bias
//                     + the already-upsampled skip image, reading the activation exactly once.
//
// Both are HBM-bound streaming kernels.  Threads of a warp cover 32 consecutive float4 channel groups of one pixel,
// so every load / store instruction moves 512 contiguous bytes.
#include "common.cuh"
#include "tc_common.cuh"

namespace rick {
namespace {

struct BlurParams {
    const float* x;
    float* out;
    float* out2;
    const float* taps;   // (4,4) as registered in Blur.kernel (already includes the upsample gain)
    const float* demod;
    const float* noise;
    const float* noise_w;
    const float* bias;
    const float* s_next;
    int batch, in_h, in_w, out_h, out_w, c4;   // c4 = C / 4
    int pad0, flip;
    int act;
    float alpha, scale;
    int strips_x, strips_y;                    // strips of TX output columns / TY output rows
};

constexpr int TX = 2;     // output columns per thread

__device__ __forceinline__ float4 f4_fma(float4 a, float w, float4 acc) {
    acc.x = fmaf(a.x, w, acc.x), acc.y = fmaf(a.y, w, acc.y), acc.z = fmaf(a.z, w, acc.z), acc.w = fmaf(a.w, w, acc.w);
    return acc;
}

// Thread = one float4 channel group x TX output columns, marching down TY output rows with a ring of 5 input rows in
// registers: while output row dy is computed from rows dy..dy+3, the loads of row dy+4 are already in flight
// (software prefetch), so each warp keeps >= (TX+3) x 512 B outstanding.  128-thread CTAs, 3 per SM.
template <int TY>   // output rows a thread marches down (16 for large problems, 4 when parallelism is short)
__global__ void __launch_bounds__(128, 3) blur_nhwc_kernel(BlurParams p) {
    __shared__ float s_taps[16];
    if (threadIdx.x < 16) {   // out[y] = sum_t in[y + t - pad0] * k[3 - t]  (true convolution, as upfirdn2d)
        const int ty = threadIdx.x >> 2, tx = threadIdx.x & 3;
        s_taps[threadIdx.x] = p.flip ? __ldg(p.taps + ty * 4 + tx) : __ldg(p.taps + (3 - ty) * 4 + (3 - tx));
    }
    __syncthreads();
    float w[4][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i >> 2][i & 3] = s_taps[i];

    // Work item = (sample, row strip, column strip, group of 32 channel-float4s).  The 4 warps of a CTA take 4 ADJACENT
    // column strips of the same channel group, so the 3 input columns neighbouring strips share are served by L1 (with
    // channel-major numbering they sat in different CTAs: 4 % L1 hit rate, 2.5x L2 traffic in the first profile).
    const int cgb_n = (p.c4 + 31) / 32;                       // 32-lane channel groups
    const int sxb_n = (p.strips_x + 3) / 4;                   // blocks of 4 column strips
    const long long total = (long long)p.batch * p.strips_y * sxb_n * cgb_n;
    const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long item = blockIdx.x; item < total; item += gridDim.x) {
        const int cgb = (int)(item % cgb_n);
        long long r = item / cgb_n;
        const int sxb = (int)(r % sxb_n);
        r /= sxb_n;
        const int sy = (int)(r % p.strips_y);
        const int b = (int)(r / p.strips_y);
        const int cg = cgb * 32 + lane;
        const int sx = sxb * 4 + warp;
        if (cg >= p.c4 || sx >= p.strips_x) continue;
        const int ox0 = sx * TX, oy0 = sy * TY;
        const int ix0 = ox0 - p.pad0, iy0 = oy0 - p.pad0;
        const float4* xin = reinterpret_cast<const float4*>(p.x) + (long long)b * p.in_h * p.in_w * p.c4 + cg;

        float4 dm = make_float4(1.f, 1.f, 1.f, 1.f), bs = make_float4(0.f, 0.f, 0.f, 0.f), sn = dm;
        if (p.demod) dm = __ldg(reinterpret_cast<const float4*>(p.demod) + (long long)b * p.c4 + cg);
        if (p.bias) bs = __ldg(reinterpret_cast<const float4*>(p.bias) + cg);
        if (p.s_next) sn = __ldg(reinterpret_cast<const float4*>(p.s_next) + (long long)b * p.c4 + cg);

        bool col_ok[TX + 3];
#pragma unroll
        for (int c = 0; c < TX + 3; ++c) col_ok[c] = (ix0 + c >= 0) && (ix0 + c < p.in_w);
        const int rows_here = min(TY, p.out_h - oy0);          // output rows this thread produces

        float4 ring[5][TX + 3];
        auto load_row = [&](int rel, float4 (&dst)[TX + 3]) {  // input row iy0 + rel
            const int iy = iy0 + rel;
            const bool row_ok = iy >= 0 && iy < p.in_h && rel < rows_here + 3;
            const float4* rowp = xin + ((long long)iy * p.in_w + ix0) * p.c4;
#pragma unroll
            for (int c = 0; c < TX + 3; ++c)
                dst[c] = (row_ok && col_ok[c]) ? __ldg(rowp + (long long)c * p.c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        load_row(0, ring[0]);
        load_row(1, ring[1]);
        load_row(2, ring[2]);
        load_row(3, ring[3]);
#pragma unroll
        for (int dy = 0; dy < TY; ++dy) {
            if (dy >= rows_here) break;
            load_row(dy + 4, ring[(dy + 4) % 5]);              // prefetch: consumed by the NEXT iteration
            const int oy = oy0 + dy;
#pragma unroll
            for (int ox = 0; ox < TX; ++ox) {
                if (ox0 + ox >= p.out_w) continue;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int ty = 0; ty < 4; ++ty)
#pragma unroll
                    for (int tx = 0; tx < 4; ++tx) acc = f4_fma(ring[(dy + ty) % 5][ox + tx], w[ty][tx], acc);
                const long long pix = ((long long)b * p.out_h + oy) * p.out_w + (ox0 + ox);
                const float nz = p.noise ? nw * __ldg(p.noise + pix) : 0.f;
                float4 v;
                v.x = fmaf(acc.x, dm.x, nz) + bs.x, v.y = fmaf(acc.y, dm.y, nz) + bs.y;
                v.z = fmaf(acc.z, dm.z, nz) + bs.z, v.w = fmaf(acc.w, dm.w, nz) + bs.w;
                if (p.act) {
                    v.x = (v.x > 0.f ? v.x : v.x * p.alpha) * p.scale, v.y = (v.y > 0.f ? v.y : v.y * p.alpha) * p.scale;
                    v.z = (v.z > 0.f ? v.z : v.z * p.alpha) * p.scale, v.w = (v.w > 0.f ? v.w : v.w * p.alpha) * p.scale;
                }
                const float4 vm = make_float4(v.x * sn.x, v.y * sn.y, v.z * sn.z, v.w * sn.w);
                if (p.out2) {
                    st_stream_f4(reinterpret_cast<float4*>(p.out) + pix * p.c4 + cg, v);
                    st_stream_f4(reinterpret_cast<float4*>(p.out2) + pix * p.c4 + cg, vm);
                } else {
                    st_stream_f4(reinterpret_cast<float4*>(p.out) + pix * p.c4 + cg, vm);   // sn == 1 unless only the modulated copy is wanted
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// streaming variant for large maps: TMA-staged rows, partial sums in registers, packed fp32 FMAs
// ---------------------------------------------------------------------------------------------------------------
// The register-ring kernel above tops out near half the HBM rate on the 64..256-px layers: ~64 scalar FMAs per
// float4 output already fill the issue slots a memory-bound kernel can spare, and 162 registers leave 12 warps per SM
// to cover DRAM latency.  This version changes both:
//   * A CTA owns (sample, 32 channels, 32 output columns, a segment of output rows) and streams the input rows it
//     needs through a 4-stage ring of TMA boxes (32 ch x 35 columns x 4 rows, 17.5 KB each).  The tensor map zero-fills
//     everything outside the image, so padding costs no predicate; the copies run ahead of the math without holding
//     registers, which is what hides the DRAM latency.  Halo re-reads: 3 of 35 columns, 3 rows per segment.
//   * Nothing but the current input row is kept: each row is scattered into the 4 output rows it touches (partial
//     sums for 2 columns x 4 pending rows), and a row is stored when its fourth input row has passed.
//   * The FMAs are issued as FFMA2 (two channels per instruction): 32 instead of 64 per float4 output.
// Thread = (column pair, float4 channel group): 16 x 8 = 128 threads; every LDS.128 of a warp reads four full 128-byte
// pixel rows (no bank conflicts), every STG.128 writes four.
constexpr int SB_TW = 32, SB_RS = 4, SB_NS = 4, SB_CB = 32, SB_BOXW = SB_TW + 3;
constexpr int SB_STAGE_BYTES = SB_BOXW * SB_RS * SB_CB * 4;

__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <bool EPI>   // EPI: the StyledConv epilogue (demod / noise / bias / activation / next-layer modulation) is applied
__global__ void __launch_bounds__(128, 3)
blur_nhwc_stream_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ BlurParams p, int seg_rows) {
    extern __shared__ unsigned char sb_raw[];
    __shared__ uint64_t s_full[SB_NS];
    __shared__ float s_taps[16];
    unsigned char* stages = sb_raw + ((128u - (tc::smem_u32(sb_raw) & 127u)) & 127u);   // 128-byte aligned TMA destinations

    const int tid = threadIdx.x, quad = tid & 7, cp = tid >> 3;
    const int sx = blockIdx.x % p.strips_x, cgi = blockIdx.x / p.strips_x;
    const int b = blockIdx.z;
    const int oy0 = blockIdx.y * seg_rows;
    const int rows_here = min(seg_rows, p.out_h - oy0);
    const int n_stage = (rows_here + 3 + SB_RS - 1) / SB_RS;
    const int x_in0 = sx * SB_TW - p.pad0, y_in0 = oy0 - p.pad0, c0 = cgi * SB_CB;

    if (tid < 16) {
        const int ty = tid >> 2, tx = tid & 3;
        s_taps[tid] = p.flip ? __ldg(p.taps + ty * 4 + tx) : __ldg(p.taps + (3 - ty) * 4 + (3 - tx));
    }
    if (tid == 0) {
        tc::tma_prefetch_desc(&tmap);
#pragma unroll
        for (int s = 0; s < SB_NS; ++s) tc::mbar_init(&s_full[s], 1);
        tc::fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < SB_NS && s < n_stage; ++s) {
            tc::mbar_arrive_expect_tx(&s_full[s], SB_STAGE_BYTES);
            tc::tma_load_4d(stages + s * SB_STAGE_BYTES, &tmap, &s_full[s], c0, x_in0, y_in0 + s * SB_RS, b);
        }
    }
    float2 w2[4][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) w2[i >> 2][i & 3] = make_float2(s_taps[i], s_taps[i]);

    const int cg = cgi * (SB_CB / 4) + quad;                 // float4 channel group of this thread
    const int ox0 = sx * SB_TW + 2 * cp;
    const bool ok0 = ox0 < p.out_w, ok1 = ox0 + 1 < p.out_w;
    // running pointers to this thread's two output pixels of the next completed row (column 1 = one pixel further)
    const long long pix0 = ((long long)b * p.out_h + oy0) * p.out_w + ox0;
    float4* o1 = reinterpret_cast<float4*>(p.out) + pix0 * p.c4 + cg;
    float4* o2 = EPI && p.out2 ? reinterpret_cast<float4*>(p.out2) + pix0 * p.c4 + cg : nullptr;
    const float* nzp = EPI && p.noise ? p.noise + pix0 : nullptr;
    const long long row_f4 = (long long)p.out_w * p.c4;

    float2 dm[2], bs[2], sn[2];
    dm[0] = dm[1] = sn[0] = sn[1] = make_float2(1.f, 1.f);
    bs[0] = bs[1] = make_float2(0.f, 0.f);
    float nw = 0.f;
    if (EPI) {
        if (p.demod) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.demod) + (long long)b * p.c4 + cg);
            dm[0] = make_float2(t.x, t.y), dm[1] = make_float2(t.z, t.w);
        }
        if (p.bias) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias) + cg);
            bs[0] = make_float2(t.x, t.y), bs[1] = make_float2(t.z, t.w);
        }
        if (p.s_next) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.s_next) + (long long)b * p.c4 + cg);
            sn[0] = make_float2(t.x, t.y), sn[1] = make_float2(t.z, t.w);
        }
        if (p.noise) nw = __ldg(p.noise_w);
    }

    float2 acc[2][4][2];                                     // [column][pending output row slot][channel pair]
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i >> 3][(i >> 1) & 3][i & 1] = make_float2(0.f, 0.f);

    auto finish = [&](int col, float2 lo, float2 hi) {       // epilogue + store of one float4 output
        if (!EPI) {
            st_stream_f4(o1 + (long long)col * p.c4, make_float4(lo.x, lo.y, hi.x, hi.y));
            return;
        }
        const float nz = nzp ? nw * __ldg(nzp + col) : 0.f;
        const float2 nz2 = make_float2(nz, nz);
        float2 v[2] = {__fadd2_rn(__ffma2_rn(lo, dm[0], nz2), bs[0]), __fadd2_rn(__ffma2_rn(hi, dm[1], nz2), bs[1])};
        if (p.act) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 neg = __fmul2_rn(v[h], make_float2(p.alpha, p.alpha));
                v[h] = __fmul2_rn(make_float2(v[h].x > 0.f ? v[h].x : neg.x, v[h].y > 0.f ? v[h].y : neg.y),
                                  make_float2(p.scale, p.scale));
            }
        }
        const float2 m0 = __fmul2_rn(v[0], sn[0]), m1 = __fmul2_rn(v[1], sn[1]);
        if (o2) {
            st_stream_f4(o1 + (long long)col * p.c4, make_float4(v[0].x, v[0].y, v[1].x, v[1].y));
            st_stream_f4(o2 + (long long)col * p.c4, make_float4(m0.x, m0.y, m1.x, m1.y));
        } else {
            st_stream_f4(o1 + (long long)col * p.c4, make_float4(m0.x, m0.y, m1.x, m1.y));   // sn == 1 unless only the modulated copy is wanted
        }
    };

    for (int st = 0; st < n_stage; ++st) {
        const int slot = st & (SB_NS - 1);
        tc::mbar_wait(&s_full[slot], (st / SB_NS) & 1);
        const float4* sbase = reinterpret_cast<const float4*>(stages + slot * SB_STAGE_BYTES) + (2 * cp) * (SB_CB / 4) + quad;
#pragma unroll
        for (int j = 0; j < SB_RS; ++j) {
            float4 v[5];
#pragma unroll
            for (int t = 0; t < 5; ++t) v[t] = sbase[(j * SB_BOXW + t) * (SB_CB / 4)];
#pragma unroll
            for (int col = 0; col < 2; ++col)
#pragma unroll
                for (int ty = 0; ty < 4; ++ty) {
                    const int s = (j - ty) & 3;              // input row 4*st + j is tap row ty of output row 4*st + j - ty
#pragma unroll
                    for (int tx = 0; tx < 4; ++tx) {
                        acc[col][s][0] = f2_fma(make_float2(v[col + tx].x, v[col + tx].y), w2[ty][tx], acc[col][s][0]);
                        acc[col][s][1] = f2_fma(make_float2(v[col + tx].z, v[col + tx].w), w2[ty][tx], acc[col][s][1]);
                    }
                }
            const int done = (j + 1) & 3;                    // the slot that just received its tap row 3
            const int o = st * SB_RS + j - 3;
            if (o >= 0 && o < rows_here) {
                if (ok0) finish(0, acc[0][done][0], acc[0][done][1]);
                if (ok1) finish(1, acc[1][done][0], acc[1][done][1]);
                o1 += row_f4;
                if (EPI) {
                    if (o2) o2 += row_f4;
                    if (nzp) nzp += p.out_w;
                }
            }
            acc[0][done][0] = acc[0][done][1] = acc[1][done][0] = acc[1][done][1] = make_float2(0.f, 0.f);
        }
        __syncthreads();                                     // every thread is done reading this slot
        if (tid == 0 && st + SB_NS < n_stage) {
            tc::mbar_arrive_expect_tx(&s_full[slot], SB_STAGE_BYTES);
            tc::tma_load_4d(stages + slot * SB_STAGE_BYTES, &tmap, &s_full[slot], c0, x_in0, y_in0 + (st + SB_NS) * SB_RS, b);
        }
    }
}

// one warp per pixel: 3 dot products over C channels, warp-shuffle reduction
__global__ void __launch_bounds__(256) to_rgb_nhwc_kernel(float* __restrict__ rgb, const float* __restrict__ y,
                                                          const float* __restrict__ wmod,
                                                          const float* __restrict__ bias,
                                                          const float* __restrict__ skip, int batch, int hw, int c4) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long pixels = (long long)batch * hw;
    for (long long pix = warp0; pix < pixels; pix += nwarps) {
        const int b = (int)(pix / hw);
        const int q = (int)(pix - (long long)b * hw);
        const float4* a = reinterpret_cast<const float4*>(y) + pix * c4;
        const float4* w0 = reinterpret_cast<const float4*>(wmod) + (long long)b * 3 * c4;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
        for (int i = lane; i < c4; i += 32) {
            const float4 v = ld_stream_f4(a + i);
            const float4 u0 = __ldg(w0 + i), u1 = __ldg(w0 + c4 + i), u2 = __ldg(w0 + 2 * c4 + i);
            r0 += v.x * u0.x + v.y * u0.y + v.z * u0.z + v.w * u0.w;
            r1 += v.x * u1.x + v.y * u1.y + v.z * u1.z + v.w * u1.w;
            r2 += v.x * u2.x + v.y * u2.y + v.z * u2.z + v.w * u2.w;
        }
        r0 = warp_sum(r0), r1 = warp_sum(r1), r2 = warp_sum(r2);
        if (lane < 3) {
            float v = lane == 0 ? r0 : (lane == 1 ? r1 : r2);
            const long long o = ((long long)b * 3 + lane) * hw + q;     // NCHW image
            v += __ldg(bias + lane);
            if (skip) v += __ldg(skip + o);
            rgb[o] = v;
        }
    }
}

// Four pixels per warp pass.  The single-pixel kernel above spends 15 shuffles + 15 adds per pixel on its three
// warp-wide sums -- more issue slots than a 512-byte pixel (C = 128) can pay for at HBM speed (0.32 of the copy rate in the
// round-1 profile).  Here a lane accumulates 4 pixels x 3 colours = 12 partial sums (16 slots) and the warp reduces all
// of them with ONE transposing butterfly: every exchange halves the slots a lane still carries (8 + 4 + 2 + 1 + 1 = 16
// shuffles for four pixels instead of 60), and 4 independent 16-byte loads per lane are in flight per channel group.
__global__ void __launch_bounds__(256) to_rgb_nhwc_kernel4(float* __restrict__ rgb, const float* __restrict__ y,
                                                           const float* __restrict__ wmod,
                                                           const float* __restrict__ bias,
                                                           const float* __restrict__ skip, int batch, int hw, int c4) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long quads = (long long)batch * hw / 4;            // hw % 4 == 0: the 4 pixels share their sample
    for (long long qd = warp0; qd < quads; qd += nwarps) {
        const long long pix = qd * 4;
        const int b = (int)(pix / hw);
        const int q = (int)(pix - (long long)b * hw);
        const float4* a = reinterpret_cast<const float4*>(y) + pix * c4;
        const float4* w0 = reinterpret_cast<const float4*>(wmod) + (long long)b * 3 * c4;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
        for (int i = lane; i < c4; i += 32) {
            float4 x[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) x[p] = ld_stream_f4(a + (long long)p * c4 + i);
            const float4 u0 = __ldg(w0 + i), u1 = __ldg(w0 + c4 + i), u2 = __ldg(w0 + 2 * c4 + i);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                v[p * 4 + 0] += x[p].x * u0.x + x[p].y * u0.y + x[p].z * u0.z + x[p].w * u0.w;
                v[p * 4 + 1] += x[p].x * u1.x + x[p].y * u1.y + x[p].z * u1.z + x[p].w * u1.w;
                v[p * 4 + 2] += x[p].x * u2.x + x[p].y * u2.y + x[p].z * u2.z + x[p].w * u2.w;
            }
        }
        // transposing butterfly: after the exchange over lane bit k a lane keeps half of its slots
#pragma unroll
        for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
            const bool upper = (lane & bit) != 0;
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const float send = upper ? v[j] : v[j + half];
                const float keep = upper ? v[j + half] : v[j];
                v[j] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
        }
        float r = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
        const int slot = lane >> 1, p = slot >> 2, k = slot & 3;      // slot = 4 * pixel + colour
        if ((lane & 1) == 0 && k < 3) {
            const long long o = ((long long)b * 3 + k) * hw + q + p;     // NCHW image
            r += __ldg(bias + k);
            if (skip) r += __ldg(skip + o);
            rgb[o] = r;
        }
    }
}

}  // namespace
}  // namespace rick

extern "C" int rick_blur_nhwc(void* out, const void* x, const float* taps, int batch, int in_h, int in_w, int channels,
                              int pad0, int pad1, int flip_taps, const rick_conv_epilogue* e, rick_stream_t stream) {
    using namespace rick;
    if (!out || !x || !taps || batch < 1 || in_h < 1 || in_w < 1 || channels < 4) return RICK_ERR_INVALID_ARGUMENT;
    if (channels % 4 != 0) return RICK_ERR_UNSUPPORTED;
    if (!aligned_to(out, 16) || !aligned_to(x, 16)) return RICK_ERR_ALIGNMENT;
    BlurParams p{};
    p.x = static_cast<const float*>(x), p.out = static_cast<float*>(out), p.taps = taps;
    p.batch = batch, p.in_h = in_h, p.in_w = in_w, p.c4 = channels / 4, p.pad0 = pad0, p.flip = flip_taps ? 1 : 0;
    p.out_h = in_h + pad0 + pad1 - 4 + 1, p.out_w = in_w + pad0 + pad1 - 4 + 1;
    if (p.out_h < 1 || p.out_w < 1) return RICK_ERR_INVALID_ARGUMENT;
    if (e) {
        p.out2 = static_cast<float*>(e->out2), p.demod = e->demod, p.noise = e->noise, p.noise_w = e->noise_weight;
        p.bias = e->bias, p.s_next = e->s_next, p.act = e->act, p.alpha = e->alpha, p.scale = e->scale;
        if (p.noise && !p.noise_w) return RICK_ERR_INVALID_ARGUMENT;
        if (p.out2 && !p.s_next) return RICK_ERR_INVALID_ARGUMENT;
        if ((p.demod && !aligned_to(p.demod, 16)) || (p.bias && !aligned_to(p.bias, 16)) ||
            (p.s_next && !aligned_to(p.s_next, 16)) || (p.out2 && !aligned_to(p.out2, 16)))
            return RICK_ERR_ALIGNMENT;
    }
    // ---- large maps: the TMA-streamed kernel ----
    if (p.out_w >= 48 && p.out_h >= 16 && channels % SB_CB == 0 && batch <= 65535) {
        EncodeTiledFn encode = get_encode_tiled();
        if (!encode) return RICK_ERR_UNSUPPORTED;
        CUtensorMap tmap;
        cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)in_w, (cuuint64_t)in_h, (cuuint64_t)batch};
        cuuint64_t strides[3] = {(cuuint64_t)channels * 4, (cuuint64_t)channels * in_w * 4,
                                 (cuuint64_t)channels * in_w * in_h * 4};
        cuuint32_t box[4] = {SB_CB, SB_BOXW, SB_RS, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return RICK_ERR_INVALID_ARGUMENT;
        p.strips_x = (int)ceil_div(p.out_w, SB_TW);
        const int ncg = channels / SB_CB;
        // output rows per CTA: 4k + 1 so that rows + 3 fills whole 4-row stages; shorter segments when CTAs are scarce
        int seg = 61;
        for (int cand : {61, 33, 17}) {
            seg = cand;
            if ((long long)p.strips_x * ncg * ceil_div(p.out_h, cand) * batch >= (long long)kNumSMs * 6) break;
        }
        const int nseg = (int)ceil_div(p.out_h, seg);
        if ((long long)p.strips_x * ncg > 0x7fffffffLL || nseg > 65535) return RICK_ERR_OVERFLOW;
        const size_t smem = (size_t)SB_NS * SB_STAGE_BYTES + 128;
        const bool epi = p.demod || p.noise || p.bias || p.s_next || p.out2 || p.act;
        auto kernel = epi ? blur_nhwc_stream_kernel<true> : blur_nhwc_stream_kernel<false>;
        {   // once per device and variant; kept out of later calls so that launches can be recorded into CUDA graphs
            static bool attr_done[64][2] = {};
            int dev = 0;
            RICK_CUDA_TRY(cudaGetDevice(&dev));
            if (dev < 0 || dev >= 64 || !attr_done[dev][epi]) {
                RICK_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                if (dev >= 0 && dev < 64) attr_done[dev][epi] = true;
            }
        }
        kernel<<<dim3((unsigned)(p.strips_x * ncg), (unsigned)nseg, (unsigned)batch), 128, smem,
                 static_cast<cudaStream_t>(stream)>>>(tmap, p, seg);
        RICK_CHECK_LAUNCH();
        return RICK_OK;
    }
    // long row marches amortise the 3-row halo but serialise; use them only when there are >= 8 waves of CTAs anyway
    p.strips_x = (int)ceil_div(p.out_w, TX);
    const long long threads16 = (long long)batch * ceil_div(p.out_h, 16) * p.strips_x * p.c4;
    const int ty = threads16 >= (long long)kNumSMs * 3 * 128 * 8 ? 16 : 4;
    p.strips_y = (int)ceil_div(p.out_h, ty);
    long long blocks = (long long)batch * p.strips_y * ceil_div(p.strips_x, 4) * ceil_div(p.c4, 32);   // one item per CTA pass
    const long long cap = (long long)kNumSMs * 48;
    if (blocks > cap) blocks = cap;
    if (ty == 16) blur_nhwc_kernel<16><<<(unsigned)blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    else blur_nhwc_kernel<4><<<(unsigned)blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

extern "C" int rick_to_rgb_nhwc(float* rgb, const float* y, const float* wmod, const float* bias, const float* skip,
                                int batch, int h, int w, int channels, rick_stream_t stream) {
    using namespace rick;
    if (!rgb || !y || !wmod || !bias || batch < 1 || h < 1 || w < 1 || channels < 4) return RICK_ERR_INVALID_ARGUMENT;
    if (channels % 4 != 0) return RICK_ERR_UNSUPPORTED;
    if (!aligned_to(y, 16) || !aligned_to(wmod, 16)) return RICK_ERR_ALIGNMENT;
    const long long pixels = (long long)batch * h * w;
    long long blocks = ceil_div(pixels, 8);
    const long long cap = (long long)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    if ((h * w) % 4 == 0) {
        blocks = ceil_div(pixels / 4, 8);
        if (blocks > cap) blocks = cap;
        to_rgb_nhwc_kernel4<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(rgb, y, wmod, bias, skip,
                                                                                              batch, h * w, channels / 4);
    } else {
        to_rgb_nhwc_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(rgb, y, wmod, bias, skip,
                                                                                             batch, h * w, channels / 4);
    }
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}
