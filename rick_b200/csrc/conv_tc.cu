// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.  sm_100a only.
//
// Replaces the grouped cuDNN convolutions behind ModulatedConv2d (gan_training/models/model_probe_tune.py:243-284).
// The modulation is folded into the activations (xm = x * s[b, ci]) and the demodulation into the epilogue, so the
// batch folds into GEMM-N and no per-sample weight tensor exists:
//
//     D[co, pixel] = sum_{tap} sum_{ci} Wt[tap][co][ci] * xm[b, y*si + dy(tap), x*si + dx(tap), ci]
//     out[b, oy, ox, co] = act( D * demod[b, co] + noise_w * noise[b, oy, ox] + bias[co] ) ; out2 = out * s_next[b, co]
//
// Layout: activations NHWC fp32 (channels innermost = GEMM-K contiguous), weights [tap][Cout][Cin] fp32; operands are
// consumed as TF32 (kind::tf32), accumulated in fp32 in tensor memory.
//
// CTA = 6 warps, persistent over output tiles (grid = #SMs):
//   warp 0  TMA producer: per (tap, 32-channel block) one weight box [128 co x 32 ci] and one activation box
//           [nb x th x tw pixels x 32 ci], shifted by the tap offset; out-of-bounds coordinates are zero-filled by the
//           TMA unit, which IS the convolution padding.  128-byte swizzle, 4-stage mbarrier ring.
//   warp 1  MMA issuer: one elected thread issues 4 x tcgen05.mma (M=128, N=tile pixels, K=8) per stage into one of two
//           TMEM accumulators; tcgen05.commit releases the smem stage / publishes the accumulator.
//   warps 2-5  epilogue: tcgen05.ld (lane = output channel, column = pixel), fused demod / noise / bias / leaky-ReLU,
//           coalesced NHWC stores (a warp writes 32 consecutive channels of one pixel = 128 B per instruction).
// The transposed stride-2 convolution of the upsampling layers runs as four polyphase sub-convolutions (4/2/2/1 taps)
// whose outputs interleave in the (2H+1)x(2W+1) result.
//
// The weight operand is described by strides (rick_conv_weight), so the SAME weight memory serves the forward pass
// (GEMM-M = output channels, K-major rows: stride_k == 1) and the data gradient (GEMM-M = the layer's INPUT channels,
// which are then the contiguous index: stride_m == 1, an MN-major operand staged as 32-channel blocks with the
// 128B/32B-atom swizzle).  No transposed or re-packed copy of the weights exists anywhere.  Output-channel counts that
// are not a multiple of 128 ride on TMA's out-of-bounds zero fill (the surplus accumulator rows are never stored).
#include "common.cuh"
#include "tc_common.cuh"

namespace rick {

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

namespace {

constexpr int kStages = 4;
constexpr int kBlockM = 128;        // output channels per tile
constexpr int kBlockK = 32;         // fp32 channels per stage = 128 B = one swizzle row
constexpr int kMaxN = 256;          // pixels per tile
constexpr int kABytes = kBlockM * kBlockK * 4;
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;      // two accumulators of up to 256 columns

struct PhaseDev {
    int n_taps;
    int dy[9], dx[9], widx[9];
    int out_y0, out_x0, rows, cols;
    int tiles_y, tiles_x, tile_begin;   // tile_begin: first tile index of this phase INSIDE one sample group
    // pixel tile of this phase: tw x th pixels of nb samples (any sizes, tw*th*nb <= 256); the MMA runs over
    // n_mma = roundup16(tw*th*nb) columns, the surplus columns are never stored
    int tw, th, nb, n_mma, box_bytes;
    unsigned inv_tw, inv_twth;          // ceil(2^16 / d): exact floor(n / d) for n < 256, d <= 256
};

struct ConvDev {
    int batch, cout, out_h, out_w, in_stride, out_stride;
    int n_phases;
    PhaseDev phase[4];
    int cout_tiles, kblocks, tiles_per_group, total_tiles;   // sample group = nb samples; all phases share nb
    int a_mn;                                                // weight operand is MN-major (data-gradient calls)
    // split-K (small feature maps: fewer output tiles than SMs): every tile's (tap, channel-block) iterations are cut
    // into ksplit ranges, one CTA-visit each; raw partial sums go to plane `range` of the workspace (out points at it,
    // planes split_plane elements apart) and conv_fold_kernel adds them in a fixed order and applies the epilogue
    int ksplit;
    long long split_plane;
    float* out;
    float* out2;
    const float* demod;
    const float* noise;
    const float* noise_w;
    const float* bias;
    const float* s_next;
    int act;
    float alpha, scale;
};

struct TileCoord {
    int phase, cout0, b0, y0, x0;
};

// Tile order: cout tile fastest, then the tiles of all phases of one sample group, then the next group.  Keeping the
// (up to four) polyphase passes over the same samples adjacent in time keeps their input resident in L2 (phase-major
// order re-read the whole activation tensor from DRAM once per phase: 7x the algorithmic traffic in the first profile).
__device__ __forceinline__ TileCoord decode_tile(const ConvDev& p, int t) {
    TileCoord c;
    const int ct = t % p.cout_tiles;
    int r = t / p.cout_tiles;
    const int group = r / p.tiles_per_group;
    r -= group * p.tiles_per_group;
    int ph = 0;
    while (ph + 1 < p.n_phases && r >= p.phase[ph + 1].tile_begin) ++ph;
    r -= p.phase[ph].tile_begin;
    const PhaseDev& P = p.phase[ph];
    c.phase = ph;
    c.cout0 = ct * kBlockM;
    c.b0 = group * P.nb;
    c.y0 = (r / P.tiles_x) * P.th;
    c.x0 = (r % P.tiles_x) * P.tw;
    return c;
}

// Stage B of the epilogue for one 32-column chunk: everything per column is a table read, a handful of FP ops and one
// or two coalesced stores.  Flags that are uniform per launch are template parameters so the body stays ~12 SASS
// instructions per column (the first version spent ~70 and was bound by instruction issue / I-cache misses).
template <bool ACT, bool DUAL, bool MULTI_B>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], const int* __restrict__ offtab,
                                               const float* __restrict__ nztab, const short* __restrict__ btab,
                                               float* __restrict__ outp, long long out2_delta, const ConvDev& p, int co,
                                               bool co_ok, float bias, float alpha, float scale, float& dm, float& sn,
                                               int& cur_b) {
#pragma unroll 8
    for (int j4 = 0; j4 < 32; j4 += 4) {
        const int4 oq = *reinterpret_cast<const int4*>(offtab + j4);
        const float4 nq = *reinterpret_cast<const float4*>(nztab + j4);
        const int offs[4] = {oq.x, oq.y, oq.z, oq.w};
        const float nzs[4] = {nq.x, nq.y, nq.z, nq.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int off = co_ok ? offs[k] : -1;          // element offset of (pixel, channel 0), or -1
            if (MULTI_B) {                                 // several samples per tile (small feature maps only)
                const int b = btab[j4 + k];
                if (off >= 0 && b != cur_b) {
                    cur_b = b;
                    if (p.demod) dm = __ldg(p.demod + (size_t)b * p.cout + co) * tc::kTf32TruncationComp;
                    if (p.s_next) sn = __ldg(p.s_next + (size_t)b * p.cout + co);
                }
            }
            float r = fmaf(__uint_as_float(v[j4 + k]), dm, nzs[k]) + bias;
            if (ACT) r = (r > 0.f ? r : r * alpha) * scale;
            if (off >= 0) {
                float* dst = outp + off;
                if (DUAL) {
                    dst[0] = r;
                    dst[out2_delta] = r * sn;
                } else {
                    dst[0] = r * sn;                       // sn == 1 unless only the modulated copy is wanted
                }
            }
        }
    }
}

// Stage B for tiles of ONE sample (every large feature map): the accumulator arrives with lane = output channel and
// column = pixel, which stored directly is one 4-byte store instruction per pixel and lane plus its 64-bit address
// arithmetic and a branch -- ~15 issue slots per column on a warp that has its scheduler to itself (IPC 0.13 in the
// round-2 source-level profile: 14.7 us per 128 x 256 tile, as long as the whole main loop of a Cin = 128 layer, so
// every short-K launch was bound by its epilogue).  Here the warp applies the epilogue with lane = channel, transposes
// its 32 x 32 block through shared memory (conflict-free both ways) and stores with 8 lanes per pixel: 8 predicated
// 128-bit stores per 32 columns instead of 32 branchy 32-bit ones.
template <bool ACT, bool DUAL>
__device__ __forceinline__ void epilogue_chunk_t(const uint32_t (&v)[32], const int* __restrict__ offtab,
                                                 const float* __restrict__ nztab, float* __restrict__ stage,
                                                 float* __restrict__ outq, long long out2_delta, int lane, bool grp_ok,
                                                 float bias, float alpha, float scale, float dm, float sn,
                                                 const float4 sn4) {
#pragma unroll
    for (int j4 = 0; j4 < 32; j4 += 4) {
        const float4 nq = *reinterpret_cast<const float4*>(nztab + j4);
        const float nzs[4] = {nq.x, nq.y, nq.z, nq.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float r = fmaf(__uint_as_float(v[j4 + k]), dm, nzs[k]) + bias;
            if (ACT) r = (r > 0.f ? r : r * alpha) * scale;
            if (!DUAL) r *= sn;                            // sn == 1 unless only the modulated copy is wanted
            stage[(j4 + k) * 32 + lane] = r;
        }
    }
    __syncwarp();
    const int sub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = i * 4 + sub;
        const int off = offtab[j];
        const float4 val = *reinterpret_cast<const float4*>(stage + j * 32 + c4);
        if (off >= 0 && grp_ok) {
            float* dst = outq + off + c4;
            *reinterpret_cast<float4*>(dst) = val;
            if (DUAL)
                *reinterpret_cast<float4*>(dst + out2_delta) = make_float4(val.x * sn4.x, val.y * sn4.y, val.z * sn4.z, val.w * sn4.w);
        }
    }
    __syncwarp();                                          // the block is re-used by the next chunk
}

template <bool ACT, bool DUAL>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x0,
               const __grid_constant__ CUtensorMap tmap_x1, const __grid_constant__ CUtensorMap tmap_x2,
               const __grid_constant__ CUtensorMap tmap_x3, const __grid_constant__ ConvDev p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128-byte swizzle pattern
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = kABytes + kMaxN * kBlockK * 4;      // fixed stride keeps every tile base 1024-aligned
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    // epilogue column tables (static shared memory so that reads compile to LDS, not generic loads)
    __shared__ __align__(16) int ep_pix[2 * kMaxN];      // element offset of (pixel, channel 0) or -1
    __shared__ __align__(16) float ep_nz[2 * kMaxN];     // noise_w * noise
    __shared__ short ep_b[2 * kMaxN];                    // sample index
    __shared__ __align__(16) float ep_stage[4 * 32 * 32];  // per epilogue warp: one 32 pixel x 32 channel block in transit

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_w);
        tc::tma_prefetch_desc(&tmap_x0);
        for (int s = 0; s < kStages; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&tmem_full[s], 1);
            tc::mbar_init(&tmem_empty[s], 4);
        }
        tc::fence_mbar_init();
    }
    if (warp == 1) {
        tc::tmem_alloc(tmem_base_slot, kTmemCols);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            uint32_t it = 0;
            const uint32_t smem_base = tc::smem_u32(smem);
            const uint32_t full_base = tc::smem_u32(full_bar), empty_base = tc::smem_u32(empty_bar);
            for (int vt = blockIdx.x; vt < p.total_tiles * p.ksplit; vt += gridDim.x) {
                const int t = vt / p.ksplit, sp = vt - t * p.ksplit;
                const TileCoord c = decode_tile(p, t);
                const PhaseDev& P = p.phase[c.phase];
                const CUtensorMap* tmap_x = c.phase == 0 ? &tmap_x0 : (c.phase == 1 ? &tmap_x1 : (c.phase == 2 ? &tmap_x2 : &tmap_x3));
                // Shallow-K layers (Cin <= 256) finish a tile in ~10 us, too little for a 4-stage ring to cover the DRAM
                // latency of rows that no tile has touched yet: pull the NEXT tile's activation rows into L2 now (one
                // box per distinct dy; the dx-shifted boxes overlap it).
                if (p.ksplit == 1 && p.kblocks <= 8 && t + (int)gridDim.x < p.total_tiles) {
                    const TileCoord cn = decode_tile(p, t + gridDim.x);
                    const PhaseDev& Pn = p.phase[cn.phase];
                    const CUtensorMap* tmap_n = cn.phase == 0 ? &tmap_x0 : (cn.phase == 1 ? &tmap_x1 : (cn.phase == 2 ? &tmap_x2 : &tmap_x3));
                    for (int tap = 0; tap < Pn.n_taps; ++tap) {
                        bool seen = false;
                        for (int q = 0; q < tap; ++q) seen |= (Pn.dy[q] == Pn.dy[tap]);
                        if (seen) continue;
                        for (int kb = 0; kb < p.kblocks; ++kb)
                            tc::tma_prefetch_l2_4d(tmap_n, kb * kBlockK, cn.x0 * p.in_stride + Pn.dx[tap],
                                                   cn.y0 * p.in_stride + Pn.dy[tap], cn.b0);
                    }
                }
                const int n_k = P.n_taps * p.kblocks;
                const int k0 = n_k * sp / p.ksplit, k1 = n_k * (sp + 1) / p.ksplit;
                int tap = k0 / p.kblocks, kb = k0 - tap * p.kblocks;
                // Per stage this thread only waits, posts the byte count and issues two TMAs: everything that depends on
                // the tile or the tap is hoisted (the flat loop it replaces re-derived tap offsets, tensor map and
                // shared addresses every stage: ~100 dependent instructions = 0.27 us of the 0.41 us a stage lasts,
                // straight out of the latency budget of a 4-stage ring -- round-2 source-level profile).
                const int bx = c.x0 * p.in_stride, by = c.y0 * p.in_stride;
                const uint32_t tx_bytes = kABytes + P.box_bytes;
                const int a_mn = p.a_mn, kblocks = p.kblocks, cb = c.b0, cm = a_mn ? c.cout0 / 32 : c.cout0;
                for (int kk = k0; kk < k1; ++tap, kb = 0) {
                    const int gx = bx + P.dx[tap], gy = by + P.dy[tap], wi = P.widx[tap];
                    int kb_end = kb + (k1 - kk);
                    if (kb_end > kblocks) kb_end = kblocks;
                    for (; kb < kb_end; ++kb, ++kk, ++it) {
                        const uint32_t s = it % kStages;
                        const uint32_t a_dst = smem_base + s * stage_bytes;
                        const uint32_t full = full_base + s * 8;
                        tc::mbar_wait_u32(empty_base + s * 8, ((it / kStages) & 1) ^ 1);
                        tc::mbar_arrive_expect_tx_u32(full, tx_bytes);
                        if (!a_mn)
                            tc::tma_load_3d_u32(a_dst, &tmap_w, full, kb * kBlockK, cm, wi);
                        else            // MN-major: four blocks of 32 M-channels x 32 K-rows (4 KB each), one 4-D box
                            tc::tma_load_4d_u32(a_dst, &tmap_w, full, 0, kb * kBlockK, wi, cm);
                        tc::tma_load_4d_u32(a_dst + kABytes, tmap_x, full, kb * kBlockK, gx, gy, cb);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (one thread; descriptors are slot-affine)
        if (lane == 0) {
            uint32_t it = 0;
            uint32_t tile_n = 0;
            const uint32_t smem_base = tc::smem_u32(smem);
            const uint32_t full_base = tc::smem_u32(full_bar), empty_base = tc::smem_u32(empty_bar);
            const uint32_t tfull_base = tc::smem_u32(tmem_full), tempty_base = tc::smem_u32(tmem_empty);
            // stage slot s lives stage_bytes further on: its descriptors are those of slot 0 plus s * (stage_bytes >> 4)
            // in the 14-bit start-address field (all of shared memory fits, so no carry into the next field)
            const uint64_t a_desc0 = p.a_mn ? tc::umma_desc_mn_sw128_32b(smem_base, kBlockK * 128)
                                            : tc::umma_desc_k_sw128(smem_base);
            const uint64_t b_desc0 = tc::umma_desc_k_sw128(smem_base + kABytes);
            const uint32_t a_kstep = p.a_mn ? (1024u >> 4) : 2u;    // 8 tf32 along K: 8 rows of 128 B (MN-major) / 32 B
            for (int vt = blockIdx.x; vt < p.total_tiles * p.ksplit; vt += gridDim.x, ++tile_n) {
                const int t = vt / p.ksplit, sp = vt - t * p.ksplit;
                const TileCoord c = decode_tile(p, t);
                const int n_k = p.phase[c.phase].n_taps * p.kblocks;
                const int n_kblocks = n_k * (sp + 1) / p.ksplit - n_k * sp / p.ksplit;   // iterations of this range (>= 1)
                const uint32_t idesc = tc::umma_idesc_tf32(kBlockM, p.phase[c.phase].n_mma, p.a_mn != 0, false);
                const uint32_t acc = tile_n & 1;
                tc::mbar_wait_u32(tempty_base + acc * 8, ((tile_n >> 1) & 1) ^ 1);
                tc::tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * kMaxN;
                for (int kb = 0; kb < n_kblocks; ++kb, ++it) {
                    const uint32_t s = it % kStages;
                    tc::mbar_wait_u32(full_base + s * 8, (it / kStages) & 1);
                    tc::tc_fence_after_sync();
                    const uint64_t a_desc = a_desc0 + (uint64_t)(s * (uint32_t)(stage_bytes >> 4));
                    const uint64_t b_desc = b_desc0 + (uint64_t)(s * (uint32_t)(stage_bytes >> 4));
#pragma unroll
                    for (int k = 0; k < kBlockK / 8; ++k)
                        tc::umma_tf32_ss(d_tmem, a_desc + (uint64_t)(k * a_kstep), b_desc + 2 * k, idesc, (kb | k) != 0);
                    tc::umma_commit_u32(empty_base + s * 8);               // smem stage reusable once these MMAs retire
                    if (kb == n_kblocks - 1) tc::umma_commit_u32(tfull_base + acc * 8);   // accumulator complete
                }
            }
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
        // Per tile, stage A (overlaps the MMAs of this tile): the 128 epilogue threads decode the tile's <= 256 columns
        // once -- output pixel index (or -1), noise_w * noise, sample index -- into a double-buffered shared table.
        // Stage B then costs a broadcast LDS + FMA + coalesced 128 B stores per column: no per-column global load and
        // no per-thread index arithmetic between the tcgen05.ld and the stores.
        const int quarter = warp & 3;
        const int e = threadIdx.x - 64;
        uint32_t tile_n = 0;
        const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
        const float alpha = p.alpha, scale = p.scale;
        for (int vt = blockIdx.x; vt < p.total_tiles * p.ksplit; vt += gridDim.x, ++tile_n) {
            const int t = vt / p.ksplit, sp = vt - t * p.ksplit;
            const TileCoord c = decode_tile(p, t);
            const PhaseDev& P = p.phase[c.phase];
            const uint32_t acc = tile_n & 1;
            int* pixtab = ep_pix + acc * kMaxN;
            float* nztab = ep_nz + acc * kMaxN;
            short* btab = ep_b + acc * kMaxN;
            const int n_valid = P.tw * P.th * P.nb;
            for (int n = e; n < kMaxN; n += 128) {
                int pix = -1, bb = 0;
                float nz = 0.f;
                if (n < n_valid) {
                    const int bi = (int)((unsigned)n * P.inv_twth >> 16);            // n / (tw*th)
                    const int rem = n - bi * (P.tw * P.th);
                    const int yi = (int)((unsigned)rem * P.inv_tw >> 16);             // rem / tw
                    const int m_y = c.y0 + yi, m_x = c.x0 + (rem - yi * P.tw), b = c.b0 + bi;
                    if (m_y < P.rows && m_x < P.cols && b < p.batch) {
                        pix = (b * p.out_h + m_y * p.out_stride + P.out_y0) * p.out_w + m_x * p.out_stride + P.out_x0;
                        if (p.noise) nz = nw * __ldg(p.noise + pix);
                        bb = b;
                        pix *= p.cout;                                     // element offset (host guarantees < 2^31)
                    }
                }
                pixtab[n] = pix, nztab[n] = nz, btab[n] = (short)bb;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");

            const bool co_ok = c.cout0 + quarter * 32 + lane < p.cout;     // cout % 128 != 0: surplus rows are zero, unused
            const int co = co_ok ? c.cout0 + quarter * 32 + lane : 0;
            const float bias = p.bias ? __ldg(p.bias + co) : 0.f;
            int cur_b = c.b0 < p.batch ? c.b0 : p.batch - 1;
            float dm = (p.demod ? __ldg(p.demod + (size_t)cur_b * p.cout + co) : 1.f) * tc::kTf32TruncationComp;
            float sn = p.s_next ? __ldg(p.s_next + (size_t)cur_b * p.cout + co) : 1.f;
            tc::mbar_wait(&tmem_full[acc], (tile_n >> 1) & 1);
            tc::tc_fence_after_sync();
            const uint32_t taddr = tmem_base + acc * kMaxN + (static_cast<uint32_t>(quarter * 32) << 16);
            const long long out2_delta = DUAL ? (p.out2 - p.out) : 0;
            if (P.nb > 1) {                 // several samples per tile (small maps): per-column sample switch, direct stores
                float* const outp = p.out + co + (long long)sp * p.split_plane;
                for (int n0 = 0; n0 < n_valid; n0 += 32) {
                    uint32_t v[32];
                    tc::tmem_ld_32x32b_x32(taddr + n0, v);
                    tc::tmem_ld_wait();
                    epilogue_chunk<ACT, DUAL, true>(v, pixtab + n0, nztab + n0, btab + n0, outp, out2_delta, p, co, co_ok,
                                                    bias, alpha, scale, dm, sn, cur_b);
                }
                tc::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
            } else {
                // one sample per tile: transposed 128-bit stores; the TMEM loads run one chunk ahead of the stores and the
                // accumulator is handed back to the MMA warp as soon as its last chunk sits in registers
                const int c4 = (lane & 7) * 4;
                const bool grp_ok = c.cout0 + quarter * 32 + c4 < p.cout;
                float* const outq = p.out + c.cout0 + quarter * 32 + (long long)sp * p.split_plane;
                float4 sn4 = make_float4(1.f, 1.f, 1.f, 1.f);
                if (DUAL && grp_ok) {
                    const float* sp4 = p.s_next + (size_t)cur_b * p.cout + c.cout0 + quarter * 32 + c4;
                    sn4 = make_float4(__ldg(sp4), __ldg(sp4 + 1), __ldg(sp4 + 2), __ldg(sp4 + 3));
                }
                float* const stage = ep_stage + quarter * 1024;
                uint32_t va[32], vb[32];
                tc::tmem_ld_32x32b_x32(taddr, va);
                for (int n0 = 0; n0 < n_valid; n0 += 64) {
                    const bool more1 = n0 + 32 < n_valid, more2 = n0 + 64 < n_valid;
                    tc::tmem_ld_wait();
                    if (more1) {
                        tc::tmem_ld_32x32b_x32(taddr + n0 + 32, vb);
                    } else {
                        tc::tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
                    }
                    epilogue_chunk_t<ACT, DUAL>(va, pixtab + n0, nztab + n0, stage, outq, out2_delta, lane, grp_ok, bias,
                                                alpha, scale, dm, sn, sn4);
                    if (more1) {
                        tc::tmem_ld_wait();
                        if (more2) {
                            tc::tmem_ld_32x32b_x32(taddr + n0 + 64, va);
                        } else {
                            tc::tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
                        }
                        epilogue_chunk_t<ACT, DUAL>(vb, pixtab + n0 + 32, nztab + n0 + 32, stage, outq, out2_delta, lane,
                                                    grp_ok, bias, alpha, scale, dm, sn, sn4);
                    }
                }
            }
        }
    }

    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after_sync();
        tc::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// Split-K fold: out = epilogue( sum_s ws[s] ).  A group of sg = 2^k >= splits lanes owns 4 consecutive channels of one
// output pixel (NHWC): lane j loads the quad of partial plane j -- all planes in flight at once, one L2 round trip instead
// of splits / 4 -- and the group adds them with a fixed xor-shuffle tree (deterministic); its first lane applies the
// epilogue and stores.  The TF32 truncation compensation was applied to the partials.
template <bool ACT>
__global__ void __launch_bounds__(256)
conv_fold_kernel(float* __restrict__ out, float* __restrict__ out2, const float* __restrict__ ws, int splits, int sg_log2,
                 long long plane, int cout, int hw, const float* __restrict__ demod, const float* __restrict__ noise,
                 const float* __restrict__ noise_w, const float* __restrict__ bias, const float* __restrict__ s_next,
                 float alpha, float scale) {
    const long long quads = plane / 4;
    const int cq = cout / 4;
    const float nw = noise ? __ldg(noise_w) : 0.f;
    const int lane = threadIdx.x & 31;
    const int sl = lane & ((1 << sg_log2) - 1);                 // which partial plane this lane loads
    const int per_warp = 32 >> sg_log2;                          // quads per warp
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long qw = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * per_warp; qw < quads;
         qw += warps * per_warp) {
        const long long q = qw + (lane >> sg_log2);
        const bool valid = q < quads;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && sl < splits) a = ld_stream_f4(reinterpret_cast<const float4*>(ws + (long long)sl * plane) + q);
        for (int o = (1 << sg_log2) >> 1; o > 0; o >>= 1) {
            a.x += __shfl_xor_sync(0xffffffffu, a.x, o), a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
            a.z += __shfl_xor_sync(0xffffffffu, a.z, o), a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
        }
        if (!valid || sl != 0) continue;
        const long long pix = q / cq;
        const int co = (int)(q - pix * cq) * 4;
        const int b = (int)(pix / hw);
        if (demod) {
            const float4 d = __ldg(reinterpret_cast<const float4*>(demod + (size_t)b * cout + co));
            a.x *= d.x, a.y *= d.y, a.z *= d.z, a.w *= d.w;
        }
        const float nz = noise ? nw * __ldg(noise + pix) : 0.f;
        if (bias) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + co));
            a.x += nz + bb.x, a.y += nz + bb.y, a.z += nz + bb.z, a.w += nz + bb.w;
        } else {
            a.x += nz, a.y += nz, a.z += nz, a.w += nz;
        }
        if (ACT) {
            a.x = (a.x > 0.f ? a.x : a.x * alpha) * scale, a.y = (a.y > 0.f ? a.y : a.y * alpha) * scale;
            a.z = (a.z > 0.f ? a.z : a.z * alpha) * scale, a.w = (a.w > 0.f ? a.w : a.w * alpha) * scale;
        }
        float4 m = a;
        if (s_next) {
            const float4 sn = __ldg(reinterpret_cast<const float4*>(s_next + (size_t)b * cout + co));
            m.x *= sn.x, m.y *= sn.y, m.z *= sn.z, m.w *= sn.w;
        }
        if (out2) {
            reinterpret_cast<float4*>(out)[q] = a;
            reinterpret_cast<float4*>(out2)[q] = m;
        } else {
            reinterpret_cast<float4*>(out)[q] = m;
        }
    }
}

// Number of K ranges per output tile: 1 unless the launch would leave most SMs idle.
int choose_ksplit(const rick_conv_geom* g, int total_tiles) {
    if (total_tiles >= kNumSMs / 2) return 1;
    long long covered = 0;
    int min_nk = 1 << 30;
    for (int i = 0; i < g->n_phases; ++i) {
        covered += (long long)g->phase[i].rows * g->phase[i].cols;
        const int nk = g->phase[i].n_taps * (g->cin / kBlockK);
        if (nk < min_nk) min_nk = nk;
    }
    if (covered < (long long)g->out_h * g->out_w) return 1;      // pixels no phase writes would be garbage in the planes
    int s = kNumSMs / total_tiles;
    if (s > min_nk / 4) s = min_nk / 4;                           // at least 4 pipeline iterations per range
    if (s > 32) s = 32;
    return s < 2 ? 1 : s;
}

}  // namespace
}  // namespace rick

extern "C" int rick_conv_tc(void* out, const void* xm, const void* wt, const rick_conv_geom* g,
                            const rick_conv_epilogue* e, rick_stream_t stream) {
    if (!g) return RICK_ERR_INVALID_ARGUMENT;
    // packed (n_weight_taps, cout, cin) weights: one K-major matrix per tap
    rick_conv_weight w{wt, (int64_t)g->cin, 1, (int64_t)g->cin * g->cout};
    return rick_conv_tc_w(out, xm, &w, g, e, nullptr, 0, stream);
}

namespace rick {
namespace {
// tile plan shared by the launcher and the workspace query
struct TilePlan {
    int ptw[4], pth[4], nb, tiles_per_group, total_tiles, ksplit;
};
int plan_tiles(const rick_conv_geom* g, TilePlan& tp);
}  // namespace
}  // namespace rick

extern "C" int64_t rick_conv_tc_workspace(const rick_conv_geom* g) {
    using namespace rick;
    if (!g || g->n_phases < 1 || g->n_phases > 4 || g->cout % 32 != 0 || g->cin % kBlockK != 0) return -1;
    TilePlan tp;
    if (plan_tiles(g, tp) != RICK_OK) return -1;
    if (tp.ksplit <= 1) return 0;
    return (int64_t)tp.ksplit * g->batch * g->out_h * g->out_w * g->cout * 4;
}

extern "C" int rick_conv_tc_plan(const rick_conv_geom* g, int* tile_w, int* tile_h, int* samples_per_tile, int* ksplit,
                                 int* total_tiles) {
    using namespace rick;
    if (!g || !tile_w || !tile_h || !samples_per_tile || !ksplit || !total_tiles) return RICK_ERR_INVALID_ARGUMENT;
    if (g->n_phases < 1 || g->n_phases > 4) return RICK_ERR_INVALID_ARGUMENT;
    if (g->cout % 32 != 0 || g->cin % kBlockK != 0) return RICK_ERR_UNSUPPORTED;
    TilePlan tp;
    const int rc = plan_tiles(g, tp);
    if (rc != RICK_OK) return rc;
    for (int i = 0; i < 4; ++i) tile_w[i] = i < g->n_phases ? tp.ptw[i] : 0, tile_h[i] = i < g->n_phases ? tp.pth[i] : 0;
    *samples_per_tile = tp.nb, *ksplit = tp.ksplit, *total_tiles = tp.total_tiles;
    return RICK_OK;
}

extern "C" int rick_conv_tc_w(void* out, const void* xm, const rick_conv_weight* wd, const rick_conv_geom* g,
                              const rick_conv_epilogue* e, void* workspace, int64_t workspace_bytes,
                              rick_stream_t stream) {
    using namespace rick;
    if (!out || !xm || !wd || !wd->ptr || !g) return RICK_ERR_INVALID_ARGUMENT;
    const void* wt = wd->ptr;
    if (g->batch < 1 || g->in_h < 1 || g->in_w < 1 || g->out_h < 1 || g->out_w < 1) return RICK_ERR_INVALID_ARGUMENT;
    if (g->n_phases < 1 || g->n_phases > 4 || g->n_weight_taps < 1) return RICK_ERR_INVALID_ARGUMENT;
    if (g->cout % 32 != 0 || g->cin % kBlockK != 0) return RICK_ERR_UNSUPPORTED;
    if (g->in_stride < 1 || g->in_stride > 2 || g->out_stride < 1 || g->out_stride > 2) return RICK_ERR_UNSUPPORTED;
    // exactly one of the two channel indices of the weight must be contiguous; every other stride a multiple of 16 B
    const bool a_mn = wd->stride_m == 1 && wd->stride_k != 1;
    if (!a_mn && wd->stride_k != 1) return RICK_ERR_UNSUPPORTED;
    if ((a_mn ? wd->stride_k : wd->stride_m) % 4 != 0 || (g->n_weight_taps > 1 && wd->stride_tap % 4 != 0) ||
        wd->stride_m < 1 || wd->stride_k < 1 || wd->stride_tap < 0)
        return RICK_ERR_ALIGNMENT;
    if (!aligned_to(out, 16) || !aligned_to(xm, 16) || !aligned_to(wt, 16)) return RICK_ERR_ALIGNMENT;
    if ((long long)g->batch * g->out_h * g->out_w * g->cout > 0x7fffffffLL) return RICK_ERR_OVERFLOW;   // 32-bit offsets
    EncodeTiledFn encode = get_encode_tiled();
    if (!encode) return RICK_ERR_UNSUPPORTED;

    TilePlan tp;
    {
        const int rc = plan_tiles(g, tp);
        if (rc != RICK_OK) return rc;
    }
    const long long out_elems = (long long)g->batch * g->out_h * g->out_w * g->cout;
    int ksplit = tp.ksplit;
    if (ksplit > 1 && (!workspace || workspace_bytes < (int64_t)ksplit * out_elems * 4 || !aligned_to(workspace, 16)))
        ksplit = 1;                                            // no (or too small a) workspace: plain launch

    ConvDev p{};
    p.batch = g->batch, p.cout = g->cout, p.out_h = g->out_h, p.out_w = g->out_w;
    p.in_stride = g->in_stride, p.out_stride = g->out_stride, p.n_phases = g->n_phases;
    p.cout_tiles = (int)ceil_div(g->cout, kBlockM), p.kblocks = g->cin / kBlockK;
    p.a_mn = a_mn ? 1 : 0;
    const int nb_common = tp.nb;
    const int* ptw = tp.ptw;
    const int* pth = tp.pth;
    int tiles = 0;
    for (int i = 0; i < g->n_phases; ++i) {
        PhaseDev& P = p.phase[i];
        const rick_conv_phase& Q = g->phase[i];
        P.n_taps = Q.n_taps;
        for (int t = 0; t < Q.n_taps; ++t) {
            if (Q.widx[t] < 0 || Q.widx[t] >= g->n_weight_taps) return RICK_ERR_INVALID_ARGUMENT;
            P.dy[t] = Q.dy[t], P.dx[t] = Q.dx[t], P.widx[t] = Q.widx[t];
        }
        P.out_y0 = Q.out_y0, P.out_x0 = Q.out_x0, P.rows = Q.rows, P.cols = Q.cols;
        if ((Q.rows - 1) * g->out_stride + Q.out_y0 >= g->out_h || (Q.cols - 1) * g->out_stride + Q.out_x0 >= g->out_w ||
            Q.out_y0 < 0 || Q.out_x0 < 0)
            return RICK_ERR_INVALID_ARGUMENT;
        const int tw = ptw[i], th = pth[i], nb = nb_common;
        P.tw = tw, P.th = th, P.nb = nb;
        P.n_mma = ((tw * th * nb + 15) / 16) * 16;
        P.box_bytes = tw * th * nb * kBlockK * 4;
        P.inv_tw = (65536u + tw - 1) / tw;
        P.inv_twth = (65536u + tw * th - 1) / (tw * th);
        P.tiles_y = (int)ceil_div(Q.rows, th), P.tiles_x = (int)ceil_div(Q.cols, tw);
        P.tile_begin = tiles;
        tiles += P.tiles_y * P.tiles_x;
    }
    p.tiles_per_group = tiles;
    const long long total = (long long)tiles * ceil_div(g->batch, nb_common) * p.cout_tiles;
    if (total > 0x7fffffffLL) return RICK_ERR_OVERFLOW;
    p.total_tiles = (int)total;
    p.out = static_cast<float*>(out);
    p.ksplit = ksplit, p.split_plane = out_elems;
    if (e) {
        if (e->noise && !e->noise_weight) return RICK_ERR_INVALID_ARGUMENT;
        if (e->out2 && !e->s_next) return RICK_ERR_INVALID_ARGUMENT;
    }
    if (ksplit > 1) {
        p.out = static_cast<float*>(workspace);                // raw partial sums; the epilogue runs in conv_fold_kernel
    } else if (e) {
        p.out2 = static_cast<float*>(e->out2), p.demod = e->demod, p.noise = e->noise, p.noise_w = e->noise_weight;
        p.bias = e->bias, p.s_next = e->s_next, p.act = e->act, p.alpha = e->alpha, p.scale = e->scale;
    }

    // ---- tensor maps ----
    CUtensorMap tmap_w, tmap_x[4];
    {
        // a single-tap weight may come with any tap stride (it is never used): give the degenerate dimension a legal one
        const cuuint64_t tap_stride = g->n_weight_taps > 1 ? (cuuint64_t)wd->stride_tap * 4 : 16;
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult rc;
        if (!a_mn) {    // K-major: rows of cin (contiguous), one row per output channel
            cuuint64_t dims[3] = {(cuuint64_t)g->cin, (cuuint64_t)g->cout, (cuuint64_t)g->n_weight_taps};
            cuuint64_t strides[2] = {(cuuint64_t)wd->stride_m * 4, tap_stride};
            cuuint32_t box[3] = {kBlockK, kBlockM, 1};
            rc = encode(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(wt), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {        // MN-major: the GEMM-M channel (cout of THIS call) is contiguous, rows run over GEMM-K (cin).
            // 4-D view (32 channels, K, taps, M/32): one box delivers the four 32-channel blocks of a tile block-major
            cuuint64_t dims[4] = {32, (cuuint64_t)g->cin, (cuuint64_t)g->n_weight_taps, (cuuint64_t)(g->cout / 32)};
            cuuint64_t strides[3] = {(cuuint64_t)wd->stride_k * 4, tap_stride, 128};
            cuuint32_t box[4] = {32, kBlockK, 1, kBlockM / 32};
            cuuint32_t estr4[4] = {1, 1, 1, 1};
            rc = encode(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(wt), dims, strides, box, estr4,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (rc != CUDA_SUCCESS) return RICK_ERR_INVALID_ARGUMENT;
    }
    for (int i = 0; i < 4; ++i) {
        const PhaseDev& P = p.phase[i < g->n_phases ? i : 0];
        cuuint64_t dims[4] = {(cuuint64_t)g->cin, (cuuint64_t)g->in_w, (cuuint64_t)g->in_h, (cuuint64_t)g->batch};
        cuuint64_t strides[3] = {(cuuint64_t)g->cin * 4, (cuuint64_t)g->cin * g->in_w * 4,
                                 (cuuint64_t)g->cin * g->in_w * g->in_h * 4};
        // with a traversal stride s the box spans tw*s input pixels and delivers ceil(box/s) = tw of them
        cuuint32_t box[4] = {kBlockK, (cuuint32_t)(P.tw * g->in_stride), (cuuint32_t)(P.th * g->in_stride),
                             (cuuint32_t)P.nb};
        cuuint32_t estr[4] = {1, (cuuint32_t)g->in_stride, (cuuint32_t)g->in_stride, 1};
        if (encode(&tmap_x[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(xm), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return RICK_ERR_INVALID_ARGUMENT;
    }

    const int stage_bytes = kABytes + kMaxN * kBlockK * 4;
    const size_t smem = 1024 + (size_t)kStages * stage_bytes + 256;
    const bool act = p.act != 0, dual = p.out2 != nullptr;
    auto kernel = act ? (dual ? conv_tc_kernel<true, true> : conv_tc_kernel<true, false>)
                      : (dual ? conv_tc_kernel<false, true> : conv_tc_kernel<false, false>);
    {   // once per device and variant; kept out of later calls so that launches can be recorded into CUDA graphs
        static bool attr_done[64][4] = {};
        int dev = 0;
        RICK_CUDA_TRY(cudaGetDevice(&dev));
        const int variant = (act ? 2 : 0) + (dual ? 1 : 0);
        if (dev < 0 || dev >= 64 || !attr_done[dev][variant]) {
            RICK_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) attr_done[dev][variant] = true;
        }
    }
    const long long visits = (long long)p.total_tiles * ksplit;
    const int grid = visits < kNumSMs ? (int)visits : kNumSMs;
    kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(tmap_w, tmap_x[0], tmap_x[1], tmap_x[2], tmap_x[3], p);
    RICK_CHECK_LAUNCH();
    if (ksplit > 1) {
        int sg_log2 = 1;
        while ((1 << sg_log2) < ksplit) ++sg_log2;              // ksplit <= 32
        long long blocks = ceil_div((out_elems / 4) << sg_log2, 256);
        if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
        const float* ws = static_cast<const float*>(workspace);
        const int hw = g->out_h * g->out_w;
        float* o = static_cast<float*>(out);
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (e && e->act)
            conv_fold_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(o, static_cast<float*>(e->out2), ws, ksplit, sg_log2,
                                                                      out_elems, g->cout, hw, e->demod, e->noise,
                                                                      e->noise_weight, e->bias, e->s_next, e->alpha, e->scale);
        else
            conv_fold_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(o, e ? static_cast<float*>(e->out2) : nullptr, ws, ksplit,
                                                                       sg_log2, out_elems, g->cout, hw, e ? e->demod : nullptr,
                                                                       e ? e->noise : nullptr, e ? e->noise_weight : nullptr,
                                                                       e ? e->bias : nullptr, e ? e->s_next : nullptr, 0.f, 1.f);
        RICK_CHECK_LAUNCH();
    }
    return RICK_OK;
}

namespace rick {
namespace {
int plan_tiles(const rick_conv_geom* g, TilePlan& tp) {
    // ---- pixel tile per phase: tw x th pixels of nb samples, tw*th*nb <= 256, chosen to minimise the padded MMA work
    //      ceil(cols/tw) * ceil(rows/th) * ceil(batch/nb) * roundup16(tw*th*nb); ties go to the wider tile.
    //      (33x33 -> 17x11 tiles, 65x65 -> 17x13, 129x129 -> 43x5: the (2H+1)^2 outputs of the transposed conv) ----
    for (int i = 0; i < g->n_phases; ++i) {
        if (g->phase[i].n_taps < 1 || g->phase[i].n_taps > 9) return RICK_ERR_INVALID_ARGUMENT;
        if (g->phase[i].rows < 1 || g->phase[i].cols < 1) return RICK_ERR_INVALID_ARGUMENT;
    }
    // Single-phase launches are also costed by WAVES over the SMs: 512 tiles of 256 pixels on 148 SMs run as 4 rounds
    // of which the last is half empty, 592 tiles of 224 pixels (32 x 7) as 4 full, shorter ones (the batch-2 layers of
    // the G step: 1/8 less time).  Several tile heights are tried per width for that reason.
    const long long outer = ceil_div(g->cout, kBlockM);
    const bool by_waves = g->n_phases == 1;
    auto choose_tile = [&](int rows, int cols, int& tw, int& th, int& nb) {
        double best = -1.0;
        const int max_side = 256 / g->in_stride;               // TMA box extent (in input pixels) must be <= 256
        for (int kx = 1; kx <= cols; ++kx) {                   // kx tiles across, balanced widths
            const int w_ = (int)ceil_div(cols, kx);
            if (w_ > max_side || w_ > kMaxN) continue;
            if (kx > 1 && w_ == (int)ceil_div(cols, kx - 1)) continue;
            int hmax = kMaxN / w_;
            if (hmax > rows) hmax = rows;
            if (hmax > max_side) hmax = max_side;
            if (hmax < 1) continue;
            const int ky0 = (int)ceil_div(rows, hmax);
            for (int ky = ky0; ky <= ky0 + (by_waves ? 6 : 0) && ky <= rows; ++ky) {
                const int h_ = (int)ceil_div(rows, ky);        // balanced heights
                int n_ = 1;
                if (w_ >= cols && h_ >= rows) {                // whole image fits: stack samples
                    n_ = kMaxN / (w_ * h_);
                    if (n_ > g->batch) n_ = g->batch;
                    if (n_ < 1) n_ = 1;
                    n_ = (int)ceil_div(g->batch, ceil_div(g->batch, n_));
                }
                const int n_mma = ((w_ * h_ * n_ + 15) / 16) * 16;
                // padded MMA work, with a mild penalty on narrow tiles (weights are re-streamed per tile)
                const long long tiles = ceil_div(cols, w_) * ceil_div(rows, h_) * ceil_div(g->batch, n_);
                double cost = (double)tiles * (n_mma + 32.0);
                if (by_waves && tiles * outer >= kNumSMs)
                    cost = (double)ceil_div(tiles * outer, kNumSMs) * kNumSMs / (double)outer * (n_mma + 32.0);
                if (best < 0 || cost < best - 1e-9) best = cost, tw = w_, th = h_, nb = n_;
            }
        }
    };
    int nb_common = 1 << 30;
    for (int i = 0; i < g->n_phases; ++i) {
        int tw = 1, th = 1, nb = 1;
        choose_tile(g->phase[i].rows, g->phase[i].cols, tw, th, nb);
        tp.ptw[i] = tw, tp.pth[i] = th;
        if (nb < nb_common) nb_common = nb;
    }
    tp.nb = nb_common;
    int tiles = 0;
    for (int i = 0; i < g->n_phases; ++i)
        tiles += (int)(ceil_div(g->phase[i].rows, tp.pth[i]) * ceil_div(g->phase[i].cols, tp.ptw[i]));
    tp.tiles_per_group = tiles;
    const long long total = (long long)tiles * ceil_div(g->batch, nb_common) * ceil_div(g->cout, kBlockM);
    if (total > 0x7fffffffLL) return RICK_ERR_OVERFLOW;
    tp.total_tiles = (int)total;
    tp.ksplit = choose_ksplit(g, tp.total_tiles);
    return RICK_OK;
}
}  // namespace
}  // namespace rick
