// upfirdn2d for sm_100a: upsample (zero-stuff) -> FIR -> downsample, with pad / crop fused.
//
// Replaces the reference's op/upfirdn2d_kernel.cu (entry point upfirdn2d_op, :209-369) behind the same
// native signature (op/upfirdn2d.cpp:12-19).  Written from the operator's definition, not from that file:
//
//   out[oy, ox] = sum_{ty, tx}  U[oy*down + ty, ox*down + tx] * taps[kh-1-ty, kw-1-tx]
//   U = zero-stuffed (factor up), then padded (negative pad = crop) input;  U[Y, X] = in[(Y-pad_y0)/up, (X-pad_x0)/up]
//   when both divisions are exact and in range, else 0.
//
// Kernels (all minor == 1, i.e. NCHW planes, except the last):
//   upfirdn2d_direct  4x4 taps, large maps, up = 2: a thread owns an OY x (16 bytes of) output patch, pulls its input
//                     window straight from global memory (L1 shares it between neighbours), FIR in registers with the
//                     polyphase structure resolved at compile time (4 FMAs per output, not 16), 128-bit stores.
//   upfirdn2d_rows    4x4 taps, large maps, up = 1 (blur, down = 2): one cp.async.bulk per CTA stages TR full-width
//                     input rows; lane = output column sliding down the rows, packed FFMA2, no predicates in the loop.
//   upfirdn2d_tiled   4x4 taps, small maps: a CTA stages an input halo tile (several planes for tiny feature maps) in
//                     shared memory with coalesced loads and implicit zero padding; 4 x OY outputs per thread.
//   upfirdn2d_sep     5 .. 16 taps per axis, up / down in {1, 2} (the 12x12 antialiasing filter of non_leaking.py): the
//                     CTA factors the outer-product kernel itself and runs two 1-D passes through shared memory.
//   upfirdn2d_generic any taps / factors / minor: one thread per output sample, taps staged in shared memory, only
//                     the taps that land on real samples are visited.
// All are HBM-bound by design: 4 B read per input + 4 B written per output sample (fp32); what limited the earlier
// versions was instruction issue, see the notes at each kernel.
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace rick {

struct UpfirdnParams {
    const void* in;
    void* out;
    const float* taps;
    long long planes;
    int in_h, in_w, out_h, out_w, minor;
    int kh, kw, flip;
    int up_x, up_y, down_x, down_y, pad_x0, pad_y0;
    // tiled kernel only
    int qx, qy;                  // floor(pad0 / UP)
    int log_txn, log_tyn, tpn;   // thread grid inside a CTA: txn x tyn threads per plane, tpn planes
    int itw, ith, itw_pad;       // staged input tile (per plane) and its row pitch in floats
    int tiles_x, tiles_y;
    // rows kernel only
    int tr, sr, ncg, nseg, edge;   // output rows per CTA / per warp item, 32-column groups, row segments, edge columns
    long long total;               // elements in the whole input tensor
};

// --------------------------------------------------------------------------------------------------------
// generic gather kernel
// --------------------------------------------------------------------------------------------------------
constexpr int kGenericMaxTaps = 1024;   // staged in shared memory up to 32 x 32 taps
template <typename T>
__global__ void __launch_bounds__(256) upfirdn2d_generic(UpfirdnParams p) {
    // taps staged once per CTA, already in the order the loops below walk them (flip resolved here)
    __shared__ float s_taps[kGenericMaxTaps];
    const int ntaps = p.kh * p.kw;
    const bool staged = ntaps <= kGenericMaxTaps;
    if (staged) {
        for (int i = threadIdx.x; i < ntaps; i += blockDim.x) {
            const int ty = i / p.kw, tx = i - ty * p.kw;
            const int ky = p.flip ? ty : p.kh - 1 - ty, kx = p.flip ? tx : p.kw - 1 - tx;
            s_taps[i] = __ldg(p.taps + ky * p.kw + kx);
        }
        __syncthreads();
    }
    const long long total = p.planes * p.out_h * p.out_w * p.minor;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int m = (int)(i % p.minor);
        long long r = i / p.minor;
        int ox = (int)(r % p.out_w);
        r /= p.out_w;
        int oy = (int)(r % p.out_h);
        long long plane = r / p.out_h;
        const T* src = static_cast<const T*>(p.in) + plane * p.in_h * (long long)p.in_w * p.minor + m;
        // tap row ty hits a real sample iff (oy*down + ty - pad) is a non-negative multiple of up: walk only those
        // (the 12x12 antialiasing filter of non_leaking.py at up = 2 touches 36 of its 144 taps per output)
        const int by = oy * p.down_y - p.pad_y0, bx = ox * p.down_x - p.pad_x0;
        const int ty0 = by >= 0 ? (p.up_y - by % p.up_y) % p.up_y : -by;
        const int tx0 = bx >= 0 ? (p.up_x - bx % p.up_x) % p.up_x : -bx;
        // consecutive valid taps read consecutive samples: one divide per output and axis, none per tap
        const int iy0 = (by + ty0) / p.up_y, ix0 = (bx + tx0) / p.up_x;
        const int ny = ty0 < p.kh ? min((p.kh - 1 - ty0) / p.up_y + 1, p.in_h - iy0) : 0;
        const int nx = tx0 < p.kw ? min((p.kw - 1 - tx0) / p.up_x + 1, p.in_w - ix0) : 0;
        float acc = 0.f;
        for (int a = 0; a < ny; ++a) {
            const T* row = src + ((long long)(iy0 + a) * p.in_w + ix0) * p.minor;
            const int ty = ty0 + a * p.up_y;
            if (staged) {
                const float* wrow = s_taps + ty * p.kw + tx0;
                for (int b = 0; b < nx; ++b) acc = fmaf(Elem<T>::ld(row + (long long)b * p.minor), wrow[b * p.up_x], acc);
            } else {
                const int ky = p.flip ? ty : p.kh - 1 - ty;
                for (int b = 0; b < nx; ++b) {
                    const int tx = tx0 + b * p.up_x;
                    const int kx = p.flip ? tx : p.kw - 1 - tx;
                    acc = fmaf(Elem<T>::ld(row + (long long)b * p.minor), __ldg(p.taps + ky * p.kw + kx), acc);
                }
            }
        }
        Elem<T>::st(static_cast<T*>(p.out) + i, acc);
    }
}

// --------------------------------------------------------------------------------------------------------
// separable tiled kernel: long FIR taps (5 .. 16 per axis), up / down in {1, 2}
// --------------------------------------------------------------------------------------------------------
// The only long filters on the path are outer products (non_leaking.py:321-323 builds its 12 x 12 antialiasing kernel as
// ger(k, k)), so a 144-tap gather per output is 6x the arithmetic the filter needs.  Each CTA factors the taps itself
// (pivot row / column of the staged 2-D kernel, verified to ~3 ulp of the pivot: no host round trip, no workspace),
// stages the input window of a 32 x 64 output tile in shared memory with the padding zero-filled, filters along x into
// a second shared tile and along y into the output.  Both passes give a thread 4 consecutive outputs ALONG the filtered
// axis, so its input window lives in registers and is shared by the 4; the polyphase structure (which taps meet which
// window element) is resolved at compile time from <UP, DOWN, KT> and the pad phase <C>, taps sit in registers.
// Kernels that are not rank 1 take a plain 2-D loop over the same staged window.
template <int UP, int DOWN, int KT>
struct SepCfg {
    static constexpr int TW = 64, TH = 32, R = 4;
    static constexpr int ITH = ((TH - 1) * DOWN + KT - 1) / UP + 2;   // staged input rows / columns per tile
    static constexpr int ITW = ((TW - 1) * DOWN + KT - 1) / UP + 2;
    static constexpr int ITWP = ITW | 1;                              // odd pitch: lanes that walk down a column
    static constexpr int TWP = TW + 1;                                //            hit 32 different banks
    static constexpr int WIN = ((R - 1) * DOWN + KT - 1) / UP + 1;    // window of 4 consecutive outputs
    static constexpr int STEP = R * DOWN / UP;                        // window shift between neighbouring groups
    static constexpr size_t smem_bytes = (size_t)(ITH * ITWP + ITH * TWP) * sizeof(float);
};

// acc[j] = sum_w win[w] * k[C + w*UP - j*DOWN]   (taps outside [0, KT) do not exist)
template <int UP, int DOWN, int KT, int C>
__device__ __forceinline__ void sep_fir4(const float (&win)[SepCfg<UP, DOWN, KT>::WIN], const float (&k)[KT], float (&acc)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = 0.f;
#pragma unroll
    for (int w = 0; w < SepCfg<UP, DOWN, KT>::WIN; ++w) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int t = C + w * UP - j * DOWN;
            if (t >= 0 && t < KT) acc[j] = fmaf(win[w], k[t], acc[j]);
        }
    }
}

template <typename T, int UP, int DOWN, int KT, int CX, int CY>
__global__ void __launch_bounds__(256) upfirdn2d_sep(UpfirdnParams p) {
    using Cfg = SepCfg<UP, DOWN, KT>;
    extern __shared__ float sep_smem[];
    float* s_in = sep_smem;                                   // [ITH][ITWP]
    float* s_tmp = sep_smem + Cfg::ITH * Cfg::ITWP;           // [ITH][TWP]
    __shared__ float s_taps[KT * KT];                         // walking order, zero padded to KT x KT
    __shared__ float s_u[KT], s_v[KT];
    __shared__ int s_sep;
    const int tid = threadIdx.x, lane = tid & 31;

    for (int i = tid; i < KT * KT; i += 256) {
        const int ty = i / KT, tx = i - ty * KT;
        float v = 0.f;
        if (ty < p.kh && tx < p.kw) {
            const int ky = p.flip ? ty : p.kh - 1 - ty, kx = p.flip ? tx : p.kw - 1 - tx;
            v = __ldg(p.taps + ky * p.kw + kx);
        }
        s_taps[i] = v;
    }
    __syncthreads();
    if (tid < 32) {                                           // rank-1 factorisation around the largest tap
        float best = -1.f;
        int bi = 0;
        for (int i = lane; i < KT * KT; i += 32) {
            const float a = fabsf(s_taps[i]);
            if (a > best) best = a, bi = i;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) best = ob, bi = oi;
        }
        const int pr = bi / KT, pc = bi - pr * KT;
        const float piv = s_taps[bi];
        if (lane < KT) {
            s_u[lane] = s_taps[lane * KT + pc];
            s_v[lane] = best > 0.f ? s_taps[pr * KT + lane] / piv : 0.f;
        }
        __syncwarp();
        float err = 0.f;
        for (int i = lane; i < KT * KT; i += 32) err = fmaxf(err, fabsf(s_taps[i] - s_u[i / KT] * s_v[i % KT]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) err = fmaxf(err, __shfl_xor_sync(0xffffffffu, err, o));
        if (lane == 0) s_sep = err <= 4e-7f * best ? 1 : 0;
    }

    // tile of this CTA
    int b = blockIdx.x;
    const int tx_i = b % p.tiles_x;
    b /= p.tiles_x;
    const int ty_i = b % p.tiles_y;
    const long long plane = b / p.tiles_y;
    const int ox0 = tx_i * Cfg::TW, oy0 = ty_i * Cfg::TH;
    // first staged input column / row: the sample tap CX (CY) of the tile's first output lands on (exact division)
    const int ix0 = (ox0 * DOWN - p.pad_x0 + CX) / UP, iy0 = (oy0 * DOWN - p.pad_y0 + CY) / UP;
    const T* src = static_cast<const T*>(p.in) + plane * p.in_h * (long long)p.in_w;
    for (int i = tid; i < Cfg::ITH * Cfg::ITW; i += 256) {
        const int r = i / Cfg::ITW, c = i - r * Cfg::ITW;
        const int iy = iy0 + r, ix = ix0 + c;
        float v = 0.f;
        if (iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w) v = Elem<T>::ld(src + (long long)iy * p.in_w + ix);
        s_in[r * Cfg::ITWP + c] = v;
    }
    __syncthreads();
    T* dst = static_cast<T*>(p.out) + plane * p.out_h * (long long)p.out_w;

    if (s_sep) {
        float k[KT];
        // pass 1, along x: item = (row r, group of 4 outputs); lanes run down the rows
#pragma unroll
        for (int i = 0; i < KT; ++i) k[i] = s_v[i];
        for (int i = tid; i < Cfg::ITH * (Cfg::TW / 4); i += 256) {
            const int g = i / Cfg::ITH, r = i - g * Cfg::ITH;
            float win[Cfg::WIN], acc[4];
            const float* row = s_in + r * Cfg::ITWP + g * Cfg::STEP;
#pragma unroll
            for (int w = 0; w < Cfg::WIN; ++w) win[w] = (g * Cfg::STEP + w < Cfg::ITW) ? row[w] : 0.f;
            sep_fir4<UP, DOWN, KT, CX>(win, k, acc);
#pragma unroll
            for (int j = 0; j < 4; ++j) s_tmp[r * Cfg::TWP + g * 4 + j] = acc[j];
        }
        __syncthreads();
        // pass 2, along y: item = (output column, group of 4 output rows); lanes run along the columns
#pragma unroll
        for (int i = 0; i < KT; ++i) k[i] = s_u[i];
        for (int i = tid; i < Cfg::TW * (Cfg::TH / 4); i += 256) {
            const int g = i / Cfg::TW, c = i - g * Cfg::TW;
            float win[Cfg::WIN], acc[4];
            const float* col = s_tmp + (g * Cfg::STEP) * Cfg::TWP + c;
#pragma unroll
            for (int w = 0; w < Cfg::WIN; ++w) win[w] = (g * Cfg::STEP + w < Cfg::ITH) ? col[w * Cfg::TWP] : 0.f;
            sep_fir4<UP, DOWN, KT, CY>(win, k, acc);
            const int ox = ox0 + c;
            if (ox < p.out_w) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int oy = oy0 + g * 4 + j;
                    if (oy < p.out_h) Elem<T>::st(dst + (long long)oy * p.out_w + ox, acc[j]);
                }
            }
        }
    } else {
        // not an outer product: 2-D sum over the staged window (same tap / window correspondence, run-time indices)
        for (int i = tid; i < Cfg::TW * Cfg::TH; i += 256) {
            const int oyl = i / Cfg::TW, oxl = i - oyl * Cfg::TW;
            const int gy = oyl >> 2, jy = oyl & 3, gx = oxl >> 2, jx = oxl & 3;
            float acc = 0.f;
            for (int wy = 0; wy < Cfg::WIN; ++wy) {
                const int ty = CY + wy * UP - jy * DOWN, r = gy * Cfg::STEP + wy;
                if (ty < 0 || ty >= KT || r >= Cfg::ITH) continue;
                for (int wx = 0; wx < Cfg::WIN; ++wx) {
                    const int tx = CX + wx * UP - jx * DOWN, c = gx * Cfg::STEP + wx;
                    if (tx < 0 || tx >= KT || c >= Cfg::ITW) continue;
                    acc = fmaf(s_in[r * Cfg::ITWP + c], s_taps[ty * KT + tx], acc);
                }
            }
            const int oy = oy0 + oyl, ox = ox0 + oxl;
            if (oy < p.out_h && ox < p.out_w) Elem<T>::st(dst + (long long)oy * p.out_w + ox, acc);
        }
    }
}

// --------------------------------------------------------------------------------------------------------
// column-walker kernel: 4x4 taps on SMALL planes (4 .. ~64 px: every layer below 64 px, any odd size)
// --------------------------------------------------------------------------------------------------------
// The tiled kernel gives a thread a 4 x OY patch and a CTA a power-of-two grid of them: on a 33 x 33 or 17 x 17 plane a
// third of the threads have work and rows end off the 128-bit store grid (0.06 - 0.3 of HBM in the round-2 op sweep).
// Here a thread owns ONE output column of one plane and walks down it: the 4 x 4 input window lives in registers and
// takes DOWN new rows per output (4 or 8 loads, straight from global memory -- neighbouring lanes read neighbouring
// samples and the three neighbours that share each sample hit L1), 16 FMAs, one store; lanes run along (plane, column)
// pairs in linear order, so every thread has work whatever the plane's size and a warp's loads / stores are contiguous
// runs.  No shared memory, no barrier, no per-output index arithmetic.  up = 2 needs no window: 2 x 2 taps per output.
template <typename T, int UP, int DOWN>
__global__ void __launch_bounds__(256) upfirdn2d_cols(UpfirdnParams p) {
    __shared__ float s_taps[16];
    const int tid = threadIdx.x;
    if (tid < 16) {
        const int ty = tid >> 2, tx = tid & 3;
        const int ky = p.flip ? ty : 3 - ty, kx = p.flip ? tx : 3 - tx;
        s_taps[tid] = __ldg(p.taps + ky * 4 + kx);
    }
    __syncthreads();
    const long long items = p.planes * p.out_w;
    const long long in_px = (long long)p.in_h * p.in_w, out_px = (long long)p.out_h * p.out_w;
    for (long long it = blockIdx.x * 256LL + tid; it < items; it += gridDim.x * 256LL) {
        const long long plane = it / p.out_w;
        const int ox = (int)(it - plane * p.out_w);
        const T* __restrict__ src = static_cast<const T*>(p.in) + plane * in_px;
        T* __restrict__ dst = static_cast<T*>(p.out) + plane * out_px + ox;
        if (UP == 1) {
            float w[4][4];
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i >> 2][i & 3] = s_taps[i];
            const int ix0 = ox * DOWN - p.pad_x0;                  // column under tap 0
            bool okx[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) okx[b] = ix0 + b >= 0 && ix0 + b < p.in_w;
            float win[4][4];
            int iy = -p.pad_y0;                                    // row under tap 0 of output row 0
#pragma unroll
            for (int a = DOWN; a < 4; ++a) {                       // rows the first output shares with its "predecessor"
                const int y = iy + a - DOWN;
                const bool oky = y >= 0 && y < p.in_h;
#pragma unroll
                for (int b = 0; b < 4; ++b) win[a][b] = (oky && okx[b]) ? Elem<T>::ld(src + (long long)y * p.in_w + ix0 + b) : 0.f;
            }
            // (unrolled so that the loads of the next rows are issued before this row's FMAs and store retire)
#pragma unroll 4
            for (int oy = 0; oy < p.out_h; ++oy, iy += DOWN) {
#pragma unroll
                for (int a = 0; a < 4 - DOWN; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) win[a][b] = win[a + DOWN][b];
#pragma unroll
                for (int a = 4 - DOWN; a < 4; ++a) {
                    const int y = iy + a;
                    const bool oky = y >= 0 && y < p.in_h;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        win[a][b] = (oky && okx[b]) ? Elem<T>::ld(src + (long long)y * p.in_w + ix0 + b) : 0.f;
                }
                float acc = 0.f;
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc = fmaf(win[a][b], w[a][b], acc);
                Elem<T>::st(dst + (long long)oy * p.out_w, acc);
            }
        } else {
            // up = 2: taps tx0, tx0 + 2 (fixed per column) and ty0, ty0 + 2 (alternating down the column) meet samples
            const int bx = ox * DOWN - p.pad_x0;
            const int tx0 = bx & 1, ix = (bx + tx0) / 2;
            const bool okx0 = ix >= 0 && ix < p.in_w, okx1 = ix + 1 >= 0 && ix + 1 < p.in_w;
            float wc[4][2];
#pragma unroll
            for (int a = 0; a < 4; ++a) wc[a][0] = s_taps[a * 4 + tx0], wc[a][1] = s_taps[a * 4 + tx0 + 2];
#pragma unroll 4
            for (int oy = 0; oy < p.out_h; ++oy) {
                const int by = oy * DOWN - p.pad_y0;
                const int ty0 = by & 1, y = (by + ty0) / 2;
                const float w00 = ty0 ? wc[1][0] : wc[0][0], w01 = ty0 ? wc[1][1] : wc[0][1];
                const float w10 = ty0 ? wc[3][0] : wc[2][0], w11 = ty0 ? wc[3][1] : wc[2][1];
                const bool oky0 = y >= 0 && y < p.in_h, oky1 = y + 1 >= 0 && y + 1 < p.in_h;
                const T* r0 = src + (long long)y * p.in_w + ix;
                const T* r1 = r0 + p.in_w;
                float acc = 0.f;
                if (oky0 && okx0) acc = fmaf(Elem<T>::ld(r0), w00, acc);
                if (oky0 && okx1) acc = fmaf(Elem<T>::ld(r0 + 1), w01, acc);
                if (oky1 && okx0) acc = fmaf(Elem<T>::ld(r1), w10, acc);
                if (oky1 && okx1) acc = fmaf(Elem<T>::ld(r1 + 1), w11, acc);
                Elem<T>::st(dst + (long long)oy * p.out_w, acc);
            }
        }
    }
}

// --------------------------------------------------------------------------------------------------------
// tiled kernel
// --------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int cfloor_div(int a, int b) {
    return (a / b) - (((a % b) != 0 && ((a < 0) != (b < 0))) ? 1 : 0);
}
__host__ __device__ constexpr int cfloor_mod(int a, int b) { return a - cfloor_div(a, b) * b; }

template <int UP, int DOWN, int P, int OUTS>
struct Window {  // 1-D footprint of OUTS consecutive outputs (first one at a multiple of UP) through 4 taps
    static constexpr int dmin = cfloor_div(0 - P, UP);
    static constexpr int dmax = cfloor_div((OUTS - 1) * DOWN + 3 - P, UP);
    static constexpr int size = dmax - dmin + 1;
};

template <typename T, int UP, int DOWN, int PX, int PY, int OY, int OX>
__global__ void __launch_bounds__(256) upfirdn2d_tiled(UpfirdnParams p) {
    static_assert(OX % 4 == 0, "outputs per thread in x must allow 128-bit stores");
    using WX = Window<UP, DOWN, PX, OX>;
    using WY = Window<UP, DOWN, PY, OY>;
    constexpr int VEC = (UP == 2) ? 2 : 4;                       // floats per shared-memory load
    constexpr int WWP = (WX::size + VEC - 1) / VEC * VEC;        // window width rounded to the vector size
    constexpr int WH = WY::size;

    extern __shared__ __align__(16) float smem[];
    __shared__ float s_taps[16];

    const int txn = 1 << p.log_txn, tyn = 1 << p.log_tyn;
    const int tid = threadIdx.x;
    if (tid < 16) {
        int ty = tid >> 2, tx = tid & 3;
        int ky = p.flip ? ty : 3 - ty, kx = p.flip ? tx : 3 - tx;
        s_taps[tid] = __ldg(p.taps + ky * 4 + kx);
    }

    // which tile / plane group
    int b = blockIdx.x;
    const int tile_x = b % p.tiles_x;
    b /= p.tiles_x;
    const int tile_y = b % p.tiles_y;
    const long long plane0 = (long long)(b / p.tiles_y) * p.tpn;

    const int tile_ox0 = tile_x * (txn * OX);
    const int tile_oy0 = tile_y * (tyn * OY);
    const int in_x0 = tile_ox0 * DOWN / UP - p.qx + WX::dmin;
    const int in_y0 = tile_oy0 * DOWN / UP - p.qy + WY::dmin;

    // ---- stage the input tile (zero outside the image: that IS the padding) ----
    // A warp owns whole tile rows (one integer divide per row, none per element) and keeps 2 rows x 4 chunks = 8
    // independent 128-byte loads in flight before the first shared-memory store.
    {
        const int warp = tid >> 5, lane = tid & 31, nwarps = (blockDim.x + 31) >> 5;   // blockDim.x may be < 32
        const int rows = p.tpn * p.ith;
        const T* in = static_cast<const T*>(p.in);
        for (int row0 = warp; row0 < rows; row0 += 2 * nwarps) {
            long long base[2];
            bool ok[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = row0 + r * nwarps;
                const int pl = row / p.ith;
                const int gy = in_y0 + (row - pl * p.ith);
                const long long plane = plane0 + pl;
                ok[r] = row < rows && gy >= 0 && gy < p.in_h && plane < p.planes;
                base[r] = (plane * p.in_h + gy) * (long long)p.in_w + in_x0;
            }
            for (int c0 = lane; c0 < p.itw_pad; c0 += 128) {
                float v[2][4];
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = c0 + 32 * u;
                        const int gx = in_x0 + c;
                        v[r][u] = (ok[r] && c < p.itw_pad && gx >= 0 && gx < p.in_w) ? Elem<T>::ld(in + base[r] + c) : 0.f;
                    }
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int row = row0 + r * nwarps;
                    if (row < rows) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int c = c0 + 32 * u;
                            if (c < p.itw_pad) smem[row * p.itw_pad + c] = v[r][u];
                        }
                    }
                }
            }
        }
    }
    __syncthreads();

    float w[4][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i >> 2][i & 3] = s_taps[i];

    const int tx_i = tid & (txn - 1);
    const int ty_i = (tid >> p.log_txn) & (tyn - 1);
    const int pl = tid >> (p.log_txn + p.log_tyn);
    const long long plane = plane0 + pl;
    const int ox0 = tile_ox0 + tx_i * OX;
    const int oy0 = tile_oy0 + ty_i * OY;
    if (plane >= p.planes || ox0 >= p.out_w || oy0 >= p.out_h) return;

    // ---- pull this thread's input window into registers ----
    float win[WH][WWP];
    {
        const float* base = smem + (pl * p.ith + ty_i * (OY * DOWN / UP)) * p.itw_pad + tx_i * (OX * DOWN / UP);
#pragma unroll
        for (int r = 0; r < WH; ++r) {
#pragma unroll
            for (int c = 0; c < WWP; c += VEC) {
                if (VEC == 4) {
                    float4 v = *reinterpret_cast<const float4*>(base + r * p.itw_pad + c);
                    win[r][c] = v.x, win[r][c + 1] = v.y, win[r][c + 2] = v.z, win[r][c + 3] = v.w;
                } else {
                    float2 v = *reinterpret_cast<const float2*>(base + r * p.itw_pad + c);
                    win[r][c] = v.x, win[r][c + 1] = v.y;
                }
            }
        }
    }

    // ---- FIR: only taps that land on real samples are instantiated ----
    T* out = static_cast<T*>(p.out) + (plane * p.out_h + oy0) * (long long)p.out_w + ox0;
    const bool full_x = (ox0 + OX <= p.out_w);
    const bool vec_ok = full_x && ((p.out_w & 3) == 0) && (sizeof(T) == 4);
    const bool fast = vec_ok && (oy0 + OY <= p.out_h);      // whole patch inside the image: no per-row / per-column checks
#pragma unroll
    for (int oy = 0; oy < OY; ++oy) {
        if (!fast && oy0 + oy >= p.out_h) break;
        float acc[OX];
#pragma unroll
        for (int ox = 0; ox < OX; ++ox) {
            float a = 0.f;
#pragma unroll
            for (int ty = 0; ty < 4; ++ty) {
                const int ny = oy * DOWN + ty - PY;
                if (cfloor_mod(ny, UP) != 0) continue;
                const int dy = cfloor_div(ny, UP) - WY::dmin;
#pragma unroll
                for (int tx = 0; tx < 4; ++tx) {
                    const int nx = ox * DOWN + tx - PX;
                    if (cfloor_mod(nx, UP) != 0) continue;
                    const int dx = cfloor_div(nx, UP) - WX::dmin;
                    a = fmaf(win[dy][dx], w[ty][tx], a);
                }
            }
            acc[ox] = a;
        }
        T* row = out + (long long)oy * p.out_w;
        if (vec_ok) {
#pragma unroll
            for (int v4 = 0; v4 < OX; v4 += 4)
                st_stream_f4(reinterpret_cast<float4*>(row + v4),
                             make_float4(acc[v4], acc[v4 + 1], acc[v4 + 2], acc[v4 + 3]));
        } else {
#pragma unroll
            for (int ox = 0; ox < OX; ++ox)
                if (ox0 + ox < p.out_w) Elem<T>::st(row + ox, acc[ox]);
        }
    }
}

// --------------------------------------------------------------------------------------------------------
// direct kernel (no shared memory) for large maps
// --------------------------------------------------------------------------------------------------------
// A thread owns an OY x 4 output patch and pulls its input window straight from global memory (neighbouring threads
// share the samples through L1), then runs the FIR out of registers and leaves 128-bit stores.  Without the
// shared-memory staging (loader loop, barrier, LDS) the instruction count per output halves and nothing serialises a
// CTA: up = 2 went from 0.68 to 0.99 of the measured HBM copy bandwidth (round-1 ncu: the tiled kernel was bound by
// instruction issue / occupancy, 63 % issue active, DRAM 53 %).  CTA = 32 x 8 threads = 128 x (8*OY) outputs.
template <typename T, int UP, int DOWN, int PX, int PY, int OY>
__global__ void __launch_bounds__(256) upfirdn2d_direct(UpfirdnParams p) {
    constexpr int OX = 16 / (int)sizeof(T);     // 16 bytes of outputs per thread and row: 4 fp32 / 8 bf16
    using WX = Window<UP, DOWN, PX, OX>;
    using WY = Window<UP, DOWN, PY, OY>;
    constexpr int WW = WX::size, WH = WY::size;

    __shared__ float s_taps[16];
    if (threadIdx.x < 16) {
        const int ty = threadIdx.x >> 2, tx = threadIdx.x & 3;
        const int ky = p.flip ? ty : 3 - ty, kx = p.flip ? tx : 3 - tx;
        s_taps[threadIdx.x] = __ldg(p.taps + ky * 4 + kx);
    }
    __syncthreads();
    float w[4][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i >> 2][i & 3] = s_taps[i];

    int b = blockIdx.x;
    const int tile_x = b % p.tiles_x;
    b /= p.tiles_x;
    const int tile_y = b % p.tiles_y;
    const long long plane = b / p.tiles_y;
    const int ox0 = (tile_x * 32 + (threadIdx.x & 31)) * OX;
    const int oy0 = (tile_y * 8 + (threadIdx.x >> 5)) * OY;
    if (ox0 >= p.out_w || oy0 >= p.out_h) return;

    const T* in = static_cast<const T*>(p.in) + plane * p.in_h * (long long)p.in_w;
    const int ix0 = ox0 * DOWN / UP - p.qx + WX::dmin;
    const int iy0 = oy0 * DOWN / UP - p.qy + WY::dmin;
    float win[WH][WW];
    bool cok[WW];
#pragma unroll
    for (int c = 0; c < WW; ++c) cok[c] = (ix0 + c >= 0) && (ix0 + c < p.in_w);
#pragma unroll
    for (int r = 0; r < WH; ++r) {
        const int iy = iy0 + r;
        const bool rok = iy >= 0 && iy < p.in_h;
        const T* rowp = in + (long long)iy * p.in_w + ix0;
#pragma unroll
        for (int c = 0; c < WW; ++c) win[r][c] = (rok && cok[c]) ? Elem<T>::ld(rowp + c) : 0.f;
    }

    T* out = static_cast<T*>(p.out) + (plane * p.out_h + oy0) * (long long)p.out_w + ox0;
    // OX outputs leave as one 16-byte store; the caller guarantees a 16-byte aligned base
    const bool vec_ok = (ox0 + OX <= p.out_w) && ((p.out_w & (OX - 1)) == 0);
    const bool fast = vec_ok && (oy0 + OY <= p.out_h);
#pragma unroll
    for (int oy = 0; oy < OY; ++oy) {
        if (!fast && oy0 + oy >= p.out_h) break;
        float acc[OX];
#pragma unroll
        for (int ox = 0; ox < OX; ++ox) {
            float a = 0.f;
#pragma unroll
            for (int ty = 0; ty < 4; ++ty) {
                const int ny = oy * DOWN + ty - PY;
                if (cfloor_mod(ny, UP) != 0) continue;
                const int dy = cfloor_div(ny, UP) - WY::dmin;
#pragma unroll
                for (int tx = 0; tx < 4; ++tx) {
                    const int nx = ox * DOWN + tx - PX;
                    if (cfloor_mod(nx, UP) != 0) continue;
                    const int dx = cfloor_div(nx, UP) - WX::dmin;
                    a = fmaf(win[dy][dx], w[ty][tx], a);
                }
            }
            acc[ox] = a;
        }
        T* row = out + (long long)oy * p.out_w;
        if (vec_ok) {
            if constexpr (sizeof(T) == 4) {
                st_stream_f4(reinterpret_cast<float4*>(row), make_float4(acc[0], acc[1], acc[2], acc[3]));
            } else {
                float4 pk;                                   // 8 bf16 = 16 bytes
                __nv_bfloat162 h2;
                h2 = __floats2bfloat162_rn(acc[0], acc[1]), pk.x = *reinterpret_cast<const float*>(&h2);
                h2 = __floats2bfloat162_rn(acc[2], acc[3]), pk.y = *reinterpret_cast<const float*>(&h2);
                h2 = __floats2bfloat162_rn(acc[4], acc[5]), pk.z = *reinterpret_cast<const float*>(&h2);
                h2 = __floats2bfloat162_rn(acc[6], acc[7]), pk.w = *reinterpret_cast<const float*>(&h2);
                st_stream_f4(reinterpret_cast<float4*>(row), pk);
            }
        } else {
#pragma unroll
            for (int ox = 0; ox < OX; ++ox)
                if (ox0 + ox < p.out_w) Elem<T>::st(row + ox, acc[ox]);
        }
    }
}

// --------------------------------------------------------------------------------------------------------
// rows kernel: up = 1 (blur / decimating blur) on large planes, staged through shared memory by one bulk copy
// --------------------------------------------------------------------------------------------------------
// For up = 1 every output needs a 4x4 window of distinct inputs, so the direct kernel spends most of its time on
// strided, mis-aligned window loads (odd widths 2^k + 1 are the common case: the blur after a transposed conv reads
// 129 -> 128, D's blur writes 128 -> 129).  Here a CTA owns TR full-width output rows of one plane: the input rows
// they need are ONE contiguous span of the NCHW tensor, fetched with a single cp.async.bulk (aligned down/up to
// 16 B; rows outside the image are not fetched, the CTA zero-fills them).  Several CTAs are resident per SM, so
// their bulk copies overlap the FIR of the others without holding registers.  Compute: lane = output column, a warp
// slides down its rows keeping the 4x4 window in registers (4 conflict-free LDS + 8 FFMA2 per output for down = 1),
// and every warp-wide store is 128 contiguous bytes.  The <= 4 columns left over when out_w = 32 k + r are reduced
// with shuffles (see "edge columns"), so 129-wide planes do not pay for a fifth, almost empty column group.
//
// The kernel is bound by instruction issue, not by bandwidth; what took it from 0.55 to 1.0 of the HBM copy rate
// (blur 129 -> 128, ncu: 70 % issue-slot utilisation, "not selected" the top stall) was removing instructions per
// output: column masks folded into the taps (no predicates), FFMA2, and ONE long item per warp in 4-warp CTAs so the
// per-CTA / per-item set-up (~250 + ~100 instructions per warp) is amortised over 48-64 rows.
template <typename T> __device__ __forceinline__ float smem_val(const T* p);
template <> __device__ __forceinline__ float smem_val<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float smem_val<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// Shared-memory image of a tile: element (row y, col x) of the plane sits at index  so + (y - y_lo) * in_w + x, for the
// R = TR*DOWN + 4 - DOWN rows starting at y_lo = oy0*DOWN - pad_y0, x in [-3, in_w + 3).  Rows outside the image are
// zero-filled by the CTA, so the FIR loop reads them like any other row.  Columns outside the image read whatever
// neighbours them in the span (finite data or zeroed slack) and are cancelled by a per-thread 0/1 mask folded into
// the taps once -- the inner loop carries no predicates.  (|pad_x| <= 3 guarantees every window keeps one real column.)
template <typename T, int DOWN>
__global__ void __launch_bounds__(256, 4) upfirdn2d_rows(UpfirdnParams p) {
    constexpr int A = 16 / (int)sizeof(T);   // elements per 16 bytes
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ uint64_t s_bar;
    __shared__ float s_taps[16];
    T* s_in = reinterpret_cast<T*>(s_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int tile_y = blockIdx.y;
    const long long plane = blockIdx.x;
    const int oy0 = tile_y * p.tr;
    const int oy1 = min(oy0 + p.tr, p.out_h);
    const int in_w = p.in_w;
    const int R = p.tr * DOWN + 4 - DOWN;
    const int y_lo = oy0 * DOWN - p.pad_y0;
    // rows [r0, r1) of the image intersect the tile's rows [y_lo, y_lo + R)
    const int r0 = max(y_lo, 0);
    const int r1 = min(y_lo + R, p.in_h);
    const T* in = static_cast<const T*>(p.in);

    int so = 3, d0 = 0, d1 = 0;          // origin offset; [d0, d1) = smem elements written by the bulk copy
    long long a0 = 0, a1 = 0, e1 = 0;
    if (r1 > r0) {
        const long long e0 = (plane * p.in_h + r0) * (long long)in_w;
        e1 = e0 + (long long)(r1 - r0) * in_w;
        a0 = e0 & ~(long long)(A - 1);
        a1 = (e1 + A - 1) & ~(long long)(A - 1);
        const long long cap = p.total & ~(long long)(A - 1);
        if (a1 > cap) a1 = cap;                      // never read past the tensor: the < 16 B tail goes by hand
        if (a1 < a0) a1 = a0;
        const int shift = (int)(e0 - a0);
        const int rel = (r0 - y_lo) * in_w;
        so = ((shift - rel) % A + A) % A;            // makes the bulk destination 16-byte aligned
        if (so < 3) so += A;
        d0 = so + rel - shift;
        d1 = d0 + (int)(a1 - a0);
    }
    const int nb = so + R * in_w + A;                // buffer length in elements
    const int row0_idx = so + (r0 - y_lo) * in_w, rowend_idx = so + (r1 - y_lo) * in_w;   // fetched rows' extent
    // zero everything the bulk copy does not write (padding rows, slack)
    for (int i = tid; i < d0; i += blockDim.x) s_in[i] = T(0.f);
    for (int i = d1 + tid; i < nb; i += blockDim.x) s_in[i] = T(0.f);
    if (tid < 16) {
        const int ty = tid >> 2, tx = tid & 3;
        const int ky = p.flip ? ty : 3 - ty, kx = p.flip ? tx : 3 - tx;
        s_taps[tid] = __ldg(p.taps + ky * 4 + kx);
    }
    if (tid == 0) {
        tc::mbar_init(&s_bar, 1);
        tc::fence_mbar_init();
    }
    __syncthreads();
    if (r1 > r0 && a1 < e1 && tid < (int)(e1 - a1)) s_in[d1 + tid] = in[a1 + tid];   // tail the copy may not touch
    const uint32_t bytes = (uint32_t)(d1 - d0) * (uint32_t)sizeof(T);
    if (bytes && tid == 0) {
        tc::mbar_arrive_expect_tx(&s_bar, bytes);
        tc::bulk_load_1d(s_in + d0, in + a0, bytes, &s_bar);
    }
    // One warp polls the mbarrier (a polling warp burns issue slots the resident CTAs' FIR loops need); the rest
    // park on the hardware barrier.  The aligned copy also brought a few elements of the rows before r0 / after
    // r1 - 1: where those rows are padding they must read as zero -- warp 0 repairs them before releasing the CTA.
    if (warp == 0) {
        if (bytes) tc::mbar_wait(&s_bar, 0);
        const bool fix_top = r1 > r0 && r0 > y_lo;
        const bool fix_bot = r1 > r0 && r1 < y_lo + R;
        if (fix_top && lane < row0_idx - d0) s_in[d0 + lane] = T(0.f);
        if (fix_bot && lane < d1 - rowend_idx) s_in[rowend_idx + lane] = T(0.f);
    }
    __syncthreads();

    T* outp = static_cast<T*>(p.out) + plane * p.out_h * (long long)p.out_w;

    // ---- main items: (32-column group, row segment); lane = output column ----
    for (int item = warp; item < p.ncg * p.nseg; item += nwarps) {
        const int cg = item % p.ncg, seg = item / p.ncg;
        const int ox = cg * 32 + lane;
        const int oys = oy0 + seg * p.sr;
        const int n = min(oys + p.sr, oy1) - oys;
        if (n <= 0) continue;
        const int x0 = ox * DOWN - p.pad_x0;
        const bool active = ox < p.out_w;
        // taps with the column mask folded in, paired along x for the packed fp32 FMA (FFMA2: two lanes per issue slot)
        float2 wm[4][2];
#pragma unroll
        for (int tx = 0; tx < 4; ++tx) {
            const float m = (active && x0 + tx >= 0 && x0 + tx < in_w) ? 1.f : 0.f;
#pragma unroll
            for (int ty = 0; ty < 4; ++ty) {
                if (tx & 1) wm[ty][tx >> 1].y = s_taps[ty * 4 + tx] * m;
                else wm[ty][tx >> 1].x = s_taps[ty * 4 + tx] * m;
            }
        }
        // inactive lanes (beyond out_w) stay inside the buffer by reading column 0
        const T* sp = s_in + so + (oys * DOWN - p.pad_y0 - y_lo) * in_w + (active ? x0 : 0);
        float2 win[4][2];
        auto load_row = [&](const T* q, float2 (&v)[2]) {
            v[0].x = smem_val<T>(q), v[0].y = smem_val<T>(q + 1), v[1].x = smem_val<T>(q + 2), v[1].y = smem_val<T>(q + 3);
        };
#pragma unroll
        for (int r = 0; r < 4 - DOWN; ++r) load_row(sp + r * in_w, win[r]);
        sp += (4 - DOWN) * in_w;
        T* orow = outp + (long long)oys * p.out_w + ox;
        auto step = [&](auto jc) {
            constexpr int j = decltype(jc)::value;
#pragma unroll
            for (int d = 0; d < DOWN; ++d) load_row(sp + d * in_w, win[(j * DOWN + (4 - DOWN) + d) & 3]);
            sp += DOWN * in_w;
            float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int ty = 0; ty < 4; ++ty) {
                a0 = __ffma2_rn(win[(j * DOWN + ty) & 3][0], wm[ty][0], a0);
                a1 = __ffma2_rn(win[(j * DOWN + ty) & 3][1], wm[ty][1], a1);
            }
            if (active) Elem<T>::st(orow, (a0.x + a0.y) + (a1.x + a1.y));
            orow += p.out_w;
        };
        int i = 0;
        for (; i + 4 <= n; i += 4) {
            step(std::integral_constant<int, 0>{});
            step(std::integral_constant<int, 1>{});
            step(std::integral_constant<int, 2>{});
            step(std::integral_constant<int, 3>{});
        }
        if (i < n) step(std::integral_constant<int, 0>{});
        if (i + 1 < n) step(std::integral_constant<int, 1>{});
        if (i + 2 < n) step(std::integral_constant<int, 2>{});
    }

    // ---- edge columns (out_w % 32 <= 4) ----
    // One column of TR outputs reads a 4-column strip of the tile; with a row pitch that is a multiple of 32 words
    // (in_w = 128, 256) a lane-per-row mapping puts all 32 lanes on one bank.  Instead every warp takes 8 output
    // rows of the column: lane = (row r = lane >> 2, tap column tx = lane & 3) loads the strip once (8-way conflict,
    // the minimum for 4 banks), rows are exchanged with shuffles and the 4 tap columns reduced with two more.
    {
        constexpr int SPAN = 8 * DOWN + 4 - DOWN;      // input rows behind 8 output rows
        constexpr int NV = (SPAN + 7) / 8;
        const int r = lane >> 2, tx = lane & 3;
        for (int item = warp; item < p.edge * (p.tr >> 3); item += nwarps) {
            const int ox = p.out_w - p.edge + item % p.edge;
            const int oyb = oy0 + (item / p.edge) * 8;
            if (oyb >= oy1) continue;
            const int x = ox * DOWN - p.pad_x0 + tx;
            const bool xok = x >= 0 && x < in_w;
            const T* sp = s_in + so + (oyb * DOWN - p.pad_y0 - y_lo) * in_w + (xok ? x : 0);
            float v[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) v[k] = (xok && r + 8 * k < SPAN) ? smem_val<T>(sp + (r + 8 * k) * in_w) : 0.f;
            float a = 0.f;
#pragma unroll
            for (int ty = 0; ty < 4; ++ty) {
                const int yr = r * DOWN + ty;             // input row (relative) feeding output row r through tap row ty
                const int src = ((yr & 7) << 2) | tx;
                float got = 0.f;
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    const float t = __shfl_sync(0xffffffffu, v[k], src);
                    if ((yr >> 3) == k) got = t;
                }
                a = fmaf(got, s_taps[ty * 4 + tx], a);
            }
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (tx == 0 && oyb + r < oy1) Elem<T>::st(outp + (long long)(oyb + r) * p.out_w + ox, a);
        }
    }
}

// returns RICK_OK after launching, or -1 when the shape does not fit this kernel (caller falls back)
template <typename T, int DOWN>
static int launch_rows(UpfirdnParams p, cudaStream_t stream) {
    constexpr int A = 16 / (int)sizeof(T);
    if (!aligned_to(p.in, 16)) return -1;
    // every window must keep at least one real column (see the kernel header)
    const int pad_x1 = (p.out_w - 1) * DOWN + 4 - p.pad_x0 - p.in_w;   // right padding actually consumed
    if (p.pad_x0 > 3 || pad_x1 > 3) return -1;
    // rows per CTA: as many as ~40 KB of staging holds (fewer, longer CTAs amortise the per-CTA and per-item set-up,
    // which is what bounds this kernel: it is instruction-issue limited, not bandwidth limited)
    int tr = 64;
    auto smem_for = [&](int t) { return ((size_t)(t * DOWN + 4 - DOWN) * p.in_w + 2 * A + 3) * sizeof(T); };
    while (tr > 8 && smem_for(tr) > 40 * 1024) tr >>= 1;
    if (smem_for(tr) > 44 * 1024) return -1;
    // balance the row tiles (129 rows -> 48 + 48 + 33, not 64 + 64 + 1); tr stays a multiple of 8 and never grows
    const int ntile = (int)ceil_div(p.out_h, tr);
    tr = std::min(tr, (int)ceil_div(ceil_div(p.out_h, ntile), 8) * 8);
    p.tr = tr;
    p.edge = (p.out_w % 32 <= 4) ? p.out_w % 32 : 0;
    p.ncg = p.out_w / 32 + ((p.out_w % 32 > 4) ? 1 : 0);
    if (p.ncg < 1) return -1;
    // one item per warp, >= 4 warps: the per-CTA and per-item set-up is replicated per warp, so fewer, longer warps win
    int nseg = 1;
    while (p.ncg * nseg < 4 && tr / (nseg * 2) >= 8 && tr % (nseg * 8) == 0) nseg <<= 1;
    p.nseg = nseg;
    const int nthreads = 32 * std::min(8, std::max(4, p.ncg * nseg));
    p.sr = tr / nseg;
    p.tiles_y = (int)ceil_div(p.out_h, tr);
    p.total = p.planes * p.in_h * (long long)p.in_w;
    if (p.planes > 0x7fffffffLL || p.tiles_y > 65535) return -1;
    const size_t smem = (smem_for(tr) + 15) & ~(size_t)15;
    upfirdn2d_rows<T, DOWN><<<dim3((unsigned)p.planes, (unsigned)p.tiles_y), nthreads, smem, stream>>>(p);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

template <typename T, int UP, int DOWN, int PX, int PY, int OY>
static int launch_direct(UpfirdnParams p, cudaStream_t stream) {
    p.tiles_x = (int)ceil_div(p.out_w, 32 * (16 / (int)sizeof(T)));
    p.tiles_y = (int)ceil_div(p.out_h, 8 * OY);
    const long long blocks = (long long)p.tiles_x * p.tiles_y * p.planes;
    if (blocks > 0x7fffffffLL) return RICK_ERR_OVERFLOW;
    upfirdn2d_direct<T, UP, DOWN, PX, PY, OY><<<(unsigned)blocks, 256, 0, stream>>>(p);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

static int ilog2_ceil(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

template <typename T, int UP, int DOWN, int PX, int PY, int OY, int OX = 4>
static int launch_tiled(UpfirdnParams p, cudaStream_t stream) {
    using WX = Window<UP, DOWN, PX, OX>;
    using WY = Window<UP, DOWN, PY, OY>;
    constexpr int VEC = (UP == 2) ? 2 : 4;
    constexpr int WWP = (WX::size + VEC - 1) / VEC * VEC;

    int log_txn = ilog2_ceil((int)ceil_div(p.out_w, OX));
    if (log_txn > 5) log_txn = 5;
    int log_tyn = ilog2_ceil((int)ceil_div(p.out_h, OY));
    if (log_tyn > 8 - log_txn) log_tyn = 8 - log_txn;
    if (log_txn == 5 && log_tyn > 3) log_tyn = 3;  // 128 x (8*OY) output tile for large maps
    const int txn = 1 << log_txn, tyn = 1 << log_tyn;
    p.log_txn = log_txn;
    p.log_tyn = log_tyn;
    p.ith = (tyn - 1) * (OY * DOWN / UP) + WY::size;
    p.itw = (txn - 1) * (OX * DOWN / UP) + WWP;
    p.itw_pad = (p.itw + 3) & ~3;
    const int plane_floats = p.ith * p.itw_pad;
    int tpn = 256 / (txn * tyn);
    while (tpn > 1 && (size_t)tpn * plane_floats * sizeof(float) > 40 * 1024) tpn >>= 1;
    while (tpn > 1 && (long long)(tpn >> 1) >= p.planes && txn * tyn * (tpn >> 1) >= 32) tpn >>= 1;
    p.tpn = tpn;
    p.tiles_x = (int)ceil_div(p.out_w, txn * OX);
    p.tiles_y = (int)ceil_div(p.out_h, tyn * OY);
    const long long blocks = (long long)p.tiles_x * p.tiles_y * ceil_div(p.planes, tpn);
    if (blocks > 0x7fffffffLL) return RICK_ERR_OVERFLOW;
    const size_t smem = (size_t)tpn * plane_floats * sizeof(float);
    upfirdn2d_tiled<T, UP, DOWN, PX, PY, OY, OX><<<(unsigned)blocks, txn * tyn * tpn, smem, stream>>>(p);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

template <typename T>
static int launch_generic(const UpfirdnParams& p, cudaStream_t stream) {
    const long long total = p.planes * p.out_h * p.out_w * p.minor;
    long long blocks = ceil_div(total, 256);
    const long long cap = (long long)kNumSMs * 32;
    if (blocks > cap) blocks = cap;
    upfirdn2d_generic<T><<<(unsigned)blocks, 256, 0, stream>>>(p);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

template <typename T, int UP, int DOWN, int KT, int CX, int CY>
static int launch_sep_c(UpfirdnParams p, cudaStream_t stream) {
    using Cfg = SepCfg<UP, DOWN, KT>;
    p.tiles_x = (int)ceil_div(p.out_w, Cfg::TW);
    p.tiles_y = (int)ceil_div(p.out_h, Cfg::TH);
    const long long blocks = (long long)p.tiles_x * p.tiles_y * p.planes;
    if (blocks > 0x7fffffffLL) return RICK_ERR_OVERFLOW;
    auto kernel = upfirdn2d_sep<T, UP, DOWN, KT, CX, CY>;
    static bool attr_done[64] = {};
    int dev = 0;
    RICK_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {       // once per device and instantiation (graph capture friendly)
        RICK_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    kernel<<<(unsigned)blocks, 256, Cfg::smem_bytes, stream>>>(p);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

template <typename T, int UP, int DOWN, int KT>
static int launch_sep_k(const UpfirdnParams& p, cudaStream_t stream) {
    if (UP == 1) return launch_sep_c<T, UP, DOWN, KT, 0, 0>(p, stream);
    const int cx = floor_mod(p.pad_x0, UP), cy = floor_mod(p.pad_y0, UP);     // first tap that meets a sample
    if (cx == 0 && cy == 0) return launch_sep_c<T, UP, DOWN, KT, 0, 0>(p, stream);
    if (cx == 1 && cy == 0) return launch_sep_c<T, UP, DOWN, KT, (UP > 1 ? 1 : 0), 0>(p, stream);
    if (cx == 0 && cy == 1) return launch_sep_c<T, UP, DOWN, KT, 0, (UP > 1 ? 1 : 0)>(p, stream);
    return launch_sep_c<T, UP, DOWN, KT, (UP > 1 ? 1 : 0), (UP > 1 ? 1 : 0)>(p, stream);
}

// long taps (5 .. 16 per axis) on NCHW planes with up / down in {1, 2}: the separable tiled kernel; -1 = not covered
template <typename T>
static int launch_sep(const UpfirdnParams& p, cudaStream_t stream) {
    if (p.minor != 1 || p.up_x != p.up_y || p.down_x != p.down_y) return -1;
    const int k = p.kh > p.kw ? p.kh : p.kw;
    if (k < 5 || k > 16) return -1;
    const int up = p.up_x, down = p.down_x;
    if (!((up == 1 && down == 1) || (up == 2 && down == 1) || (up == 1 && down == 2))) return -1;
    if (p.out_w < 16 || p.out_h < 8) return -1;            // tiny maps: the gather kernel has less to stage
#define RICK_SEP(UP_, DOWN_)                                                                    \
    if (up == UP_ && down == DOWN_) {                                                           \
        if (k <= 8) return launch_sep_k<T, UP_, DOWN_, 8>(p, stream);                           \
        if (k <= 12) return launch_sep_k<T, UP_, DOWN_, 12>(p, stream);                         \
        return launch_sep_k<T, UP_, DOWN_, 16>(p, stream);                                      \
    }
    RICK_SEP(1, 1)
    RICK_SEP(2, 1)
    RICK_SEP(1, 2)
#undef RICK_SEP
    return -1;
}

// column-walker kernel for 4x4 taps on small planes
template <typename T, int UP, int DOWN>
static int launch_cols(const UpfirdnParams& p, cudaStream_t stream) {
    const long long items = p.planes * p.out_w;
    long long blocks = ceil_div(items, 256);
    const long long cap = (long long)kNumSMs * 64;
    if (blocks > cap) blocks = cap;
    upfirdn2d_cols<T, UP, DOWN><<<(unsigned)blocks, 256, 0, stream>>>(p);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}

template <typename T>
static int dispatch(UpfirdnParams p, cudaStream_t stream) {
    // 128-bit output stores need a 16-byte aligned base (rows are then aligned whenever out_w % 4 == 0)
    const bool tiled_ok = p.minor == 1 && p.kh == 4 && p.kw == 4 && p.up_x == p.up_y && p.down_x == p.down_y &&
                          p.pad_x0 > -64 && p.pad_y0 > -64 && aligned_to(p.out, 16) &&
                          ((p.up_x == 1 && p.down_x == 1) || (p.up_x == 2 && p.down_x == 1) ||
                           (p.up_x == 1 && p.down_x == 2));
    if (!tiled_ok) {
        const int rc = launch_sep<T>(p, stream);
        return rc >= 0 ? rc : launch_generic<T>(p, stream);
    }
    const int up = p.up_x;
    p.qx = floor_div(p.pad_x0, up);
    p.qy = floor_div(p.pad_y0, up);
    const int px = floor_mod(p.pad_x0, up), py = floor_mod(p.pad_y0, up);
    const bool large = p.out_w >= 64 && p.out_h >= 32;   // large maps: the shared-memory-free kernel
    if (up == 1 && large) {                               // blur / decimating blur: bulk-staged full-width row tiles
        const int rc = p.down_x == 1 ? launch_rows<T, 1>(p, stream) : launch_rows<T, 2>(p, stream);
        if (rc >= 0) return rc;
    }
    if (!large && p.planes * p.out_w >= 4096) {           // small planes, enough columns to fill the GPU: column walkers
        if (up == 1 && p.down_x == 1) return launch_cols<T, 1, 1>(p, stream);
        if (up == 2 && p.down_x == 1) return launch_cols<T, 2, 1>(p, stream);
        if (up == 1 && p.down_x == 2) return launch_cols<T, 1, 2>(p, stream);
    }
    // tiny planes (<= 16 output rows, the 4 .. 8 px layers): 2 output rows per thread instead of 4 / 8, or a
    // (32, 512, 4, 4) -> 8 x 8 call runs on 128 CTAs of threads that each grind through 32 outputs
    const bool tiny = p.out_h <= 16;
    if (up == 1 && p.down_x == 1)
        return large ? launch_direct<T, 1, 1, 0, 0, 4>(p, stream)
                     : (tiny ? launch_tiled<T, 1, 1, 0, 0, 2>(p, stream) : launch_tiled<T, 1, 1, 0, 0, 4>(p, stream));
    if (up == 1 && p.down_x == 2) return large ? launch_direct<T, 1, 2, 0, 0, 4>(p, stream) : launch_tiled<T, 1, 2, 0, 0, 2>(p, stream);
    if (large) {
        if (px == 0 && py == 0) return launch_direct<T, 2, 1, 0, 0, 8>(p, stream);
        if (px == 1 && py == 0) return launch_direct<T, 2, 1, 1, 0, 8>(p, stream);
        if (px == 0 && py == 1) return launch_direct<T, 2, 1, 0, 1, 8>(p, stream);
        return launch_direct<T, 2, 1, 1, 1, 8>(p, stream);
    }
    // (an 8-wide patch per thread was measured slower: 3.94 vs 4.43 TB/s -- its two 128-bit stores per row leave every
    //  warp-wide store instruction half-covering its 32-byte sectors; 4-wide keeps each store instruction at 512
    //  contiguous bytes)
    if (tiny) {
        if (px == 0 && py == 0) return launch_tiled<T, 2, 1, 0, 0, 2>(p, stream);
        if (px == 1 && py == 0) return launch_tiled<T, 2, 1, 1, 0, 2>(p, stream);
        if (px == 0 && py == 1) return launch_tiled<T, 2, 1, 0, 1, 2>(p, stream);
        return launch_tiled<T, 2, 1, 1, 1, 2>(p, stream);
    }
    if (px == 0 && py == 0) return launch_tiled<T, 2, 1, 0, 0, 8>(p, stream);
    if (px == 1 && py == 0) return launch_tiled<T, 2, 1, 1, 0, 8>(p, stream);
    if (px == 0 && py == 1) return launch_tiled<T, 2, 1, 0, 1, 8>(p, stream);
    return launch_tiled<T, 2, 1, 1, 1, 8>(p, stream);
}

}  // namespace rick

extern "C" int rick_upfirdn2d_out_size(int in_size, int k, int up, int down, int pad0, int pad1) {
    if (up < 1 || down < 1 || k < 1) return -1;
    return rick::floor_div(in_size * up + pad0 + pad1 - k, down) + 1;
}

extern "C" int rick_upfirdn2d(void* out, const void* in, const float* taps, int64_t major, int in_h, int in_w,
                              int minor, int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0,
                              int pad_x1, int pad_y0, int pad_y1, int flip_taps, int dtype, rick_stream_t stream) {
    using namespace rick;
    if (!out || !in || !taps) return RICK_ERR_INVALID_ARGUMENT;
    if (major < 0 || in_h < 1 || in_w < 1 || minor < 1 || kh < 1 || kw < 1 || up_x < 1 || up_y < 1 || down_x < 1 ||
        down_y < 1)
        return RICK_ERR_INVALID_ARGUMENT;
    if (dtype != RICK_F32 && dtype != RICK_BF16) return RICK_ERR_INVALID_ARGUMENT;
    // a single plane must fit int32 coordinates (element offsets are 64-bit throughout)
    if ((long long)in_h * up_y + 2LL * (kh + 64) > 0x3fffffffLL || (long long)in_w * up_x + 2LL * (kw + 64) > 0x3fffffffLL)
        return RICK_ERR_OVERFLOW;
    const int out_h = rick_upfirdn2d_out_size(in_h, kh, up_y, down_y, pad_y0, pad_y1);
    const int out_w = rick_upfirdn2d_out_size(in_w, kw, up_x, down_x, pad_x0, pad_x1);
    if (out_h < 1 || out_w < 1) return RICK_ERR_INVALID_ARGUMENT;
    if (major == 0) return RICK_OK;
    const size_t esz = dtype == RICK_F32 ? 4 : 2;
    if (!aligned_to(out, esz) || !aligned_to(in, esz) || !aligned_to(taps, 4)) return RICK_ERR_ALIGNMENT;

    UpfirdnParams p{};
    p.in = in, p.out = out, p.taps = taps;
    p.planes = major, p.in_h = in_h, p.in_w = in_w, p.out_h = out_h, p.out_w = out_w, p.minor = minor;
    p.kh = kh, p.kw = kw, p.flip = flip_taps ? 1 : 0;
    p.up_x = up_x, p.up_y = up_y, p.down_x = down_x, p.down_y = down_y, p.pad_x0 = pad_x0, p.pad_y0 = pad_y0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return dtype == RICK_F32 ? dispatch<float>(p, s) : dispatch<__nv_bfloat16>(p, s);
}
