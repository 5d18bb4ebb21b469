// Weight gradient of the convolutions of rick_conv_tc on the 5th-generation tensor cores.  sm_100a only.
//
// Stands where the reference relies on autograd's cuDNN weight-gradient kernels behind F.conv2d / F.conv_transpose2d
// (gan_training/models/model_probe_tune.py:122-128, 265, 274, 280).  For every filter tap t:
//
//     dW[t][co][ci] = sum over (b, m, n)  G[b, m*gs + gy[t], n*gs + gx[t], co] * X[b, m*xs + xy[t], n*xs + xx[t], ci]
//
// i.e. per tap a GEMM with M = Cout, N = Cin and K = the pixels of the batch.  Activations are NHWC, so for BOTH operands
// the contiguous index is the GEMM-M / GEMM-N channel and the reduction index (pixels) is the strided one: both are
// MN-major UMMA operands, staged by TMA as blocks of 32 channels x 32 pixels with the 128B/32B-atom swizzle (the only
// swizzled layout 32-bit MN-major operands have; tc_common.cuh).  Out-of-range pixels are zero-filled by the TMA unit,
// which implements the convolution padding and ragged edges for free.
//
// Work decomposition: item = (group of taps, 128-channel block of Cout, <=256-channel block of Cin).  Layers with
// Cin <= 128 take three taps per item (accumulator columns = 3 x Cin <= 384): the G tile is loaded once for the three of
// them, which lifts the flops per staged byte to the level of the forward kernel (a lone 128 x 128 tile moves 32 KB per
// 1 MFLOP and starves the tensor pipe: 21 % active in the first profile).  The pixel axis of an item is split over
// `splits` CTAs so that items x splits is one full wave of the GPU.  CTA = 6 warps:
//   warp 0  TMA producer: per 32-pixel K-tile ONE 5-D box of G (4 channel blocks) and one per tap of X (Cin/32 channel
//           blocks each) into an mbarrier ring -- the tensor maps view the channel axis as (C/32, 32) so that a single
//           copy delivers all 32-channel blocks in the block-major order the MN-major descriptors expect
//   warp 1  MMA issuer: per stage 4 k-steps of tcgen05.mma.kind::tf32 (M = 128, N <= 256, K = 8; two per k-step when the
//           item has more than 256 columns) into one TMEM accumulator of up to 512 columns
//   warps 2-5  epilogue: tcgen05.ld, each thread owns one output channel row and writes 128 B runs of the partial result
// Partials land in a workspace [split][tap][Cout][Cin]; wgrad_fold adds the splits in a fixed order (deterministic) and
// writes the gradient with the caller's strides, so it arrives in the parameter's own memory layout.
#include "common.cuh"
#include "tc_common.cuh"

namespace rick {
namespace {

constexpr int kMaxStages = 4;
constexpr int kBlockM = 128;         // output channels (GEMM-M) per item
constexpr int kMaxN = 256;           // input channels per Cin block = columns of one MMA
constexpr int kMaxCols = 384;        // accumulator columns per item (taps_per_item x n_tile)
constexpr int kTileK = 32;           // pixels per pipeline stage
constexpr int kBlockBytes = kTileK * 128;     // one staged block: 32 pixels x 32 channels x 4 B
constexpr int kThreads = 192;
constexpr int kSmemBudget = 200 * 1024;

struct WgradDev {
    int batch, cout, cin, n_taps;
    int cin_tiles, n_tile;            // Cin is cut into cin_tiles blocks of n_tile channels (n_tile % 32 == 0, <= 256)
    int taps_per_item, tap_groups;    // an item accumulates taps_per_item taps side by side (columns = taps x n_tile)
    int stages, stage_bytes;
    int cout_tiles, items, splits;
    int tw, th, nb, tiles_x, tiles_y, tiles_b, k_tiles;      // pixel tile = tw x th pixels of nb samples = 32 rows
    int g_stride, x_stride;
    int gy[9], gx[9], xy[9], xx[9];
    float* ws;
};

__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                  const __grid_constant__ WgradDev p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = p.stage_bytes;                                 // multiple of 4 KB: every block 1024-aligned
    const int kStages = p.stages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full = empty_bar + kMaxStages;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- which (tap, cout block, cin block, pixel range) this CTA owns
    const int item = blockIdx.x % p.items;
    const int split = blockIdx.x / p.items;
    const int ci_t = item % p.cin_tiles;
    const int co_t = (item / p.cin_tiles) % p.cout_tiles;
    const int tap0 = (item / (p.cin_tiles * p.cout_tiles)) * p.taps_per_item;
    const int n_item_taps = min(p.taps_per_item, p.n_taps - tap0);
    const int co0 = co_t * kBlockM, ci0 = ci_t * p.n_tile;
    const int k_begin = (int)((long long)p.k_tiles * split / p.splits);
    const int k_end = (int)((long long)p.k_tiles * (split + 1) / p.splits);
    const int n_blocks = p.n_tile / 32;
    const int n_cols = n_item_taps * p.n_tile;                 // accumulator columns in use (<= 384)
    const uint32_t tmem_cols = n_cols <= 32 ? 32u : (n_cols <= 64 ? 64u : (n_cols <= 128 ? 128u : (n_cols <= 256 ? 256u : 512u)));

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_g);
        tc::tma_prefetch_desc(&tmap_x);
        for (int s = 0; s < kStages; ++s) {
            tc::mbar_init(&full_bar[s], 1);
            tc::mbar_init(&empty_bar[s], 1);
        }
        tc::mbar_init(tmem_full, 1);
        tc::fence_mbar_init();
    }
    if (warp == 1) {
        tc::tmem_alloc(tmem_base_slot, tmem_cols);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(kBlockM / 32 + n_item_taps * n_blocks) * kBlockBytes;
            uint32_t it = 0;
            for (int kt = k_begin; kt < k_end; ++kt, ++it) {
                const int tx = kt % p.tiles_x;
                const int r = kt / p.tiles_x;
                const int ty = r % p.tiles_y;
                const int tb = r / p.tiles_y;
                const int m0 = ty * p.th, n0 = tx * p.tw, b0 = tb * p.nb;
                const int s = it % kStages;
                tc::mbar_wait(&empty_bar[s], ((it / kStages) & 1) ^ 1);
                uint8_t* a_dst = smem + s * stage_bytes;
                uint8_t* b_dst = a_dst + (kBlockM / 32) * kBlockBytes;
                tc::mbar_arrive_expect_tx(&full_bar[s], bytes);
                // G does not depend on the tap for an ordinary convolution (gy = gx = 0) but does for the transposed one;
                // an item's taps share one G tile only in the former case (the host groups taps only then)
                tc::tma_load_5d(a_dst, &tmap_g, &full_bar[s], 0, n0 * p.g_stride + p.gx[tap0], m0 * p.g_stride + p.gy[tap0],
                                b0, co0 / 32);
                for (int j = 0; j < n_item_taps; ++j)
                    tc::tma_load_5d(b_dst + j * n_blocks * kBlockBytes, &tmap_x, &full_bar[s], 0,
                                    n0 * p.x_stride + p.xx[tap0 + j], m0 * p.x_stride + p.xy[tap0 + j], b0, ci0 / 32);
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        const int n_first = n_cols <= kMaxN ? n_cols : kMaxN;          // columns of the first MMA of a k-step
        const uint32_t idesc0 = tc::umma_idesc_tf32(kBlockM, n_first, true, true);
        const uint32_t idesc1 = tc::umma_idesc_tf32(kBlockM, n_cols > kMaxN ? n_cols - kMaxN : 16, true, true);
        uint32_t it = 0;
        for (int kt = k_begin; kt < k_end; ++kt, ++it) {
            const int s = it % kStages;
            tc::mbar_wait(&full_bar[s], (it / kStages) & 1);
            tc::tc_fence_after_sync();
            if (tc::elect_one()) {
                const uint32_t a_addr = tc::smem_u32(smem + s * stage_bytes);
                const uint32_t b_addr = a_addr + (kBlockM / 32) * kBlockBytes;
#pragma unroll
                for (int k = 0; k < kTileK / 8; ++k) {
                    const uint64_t a_desc = tc::umma_desc_mn_sw128_32b(a_addr + k * 1024, kBlockBytes);
                    const uint64_t b_desc = tc::umma_desc_mn_sw128_32b(b_addr + k * 1024, kBlockBytes);
                    tc::umma_tf32_ss(tmem_base, a_desc, b_desc, idesc0, (it | k) != 0);
                    if (n_cols > kMaxN) {                                   // columns 256.. : the blocks that follow
                        const uint64_t b_desc1 =
                            tc::umma_desc_mn_sw128_32b(b_addr + (kMaxN / 32) * kBlockBytes + k * 1024, kBlockBytes);
                        tc::umma_tf32_ss(tmem_base + kMaxN, a_desc, b_desc1, idesc1, (it | k) != 0);
                    }
                }
                tc::umma_commit(&empty_bar[s]);
                if (kt == k_end - 1) tc::umma_commit(tmem_full);
            }
            __syncwarp();
        }
    } else {
        // ===================================================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
        const int quarter = warp & 3;
        const int co = co0 + quarter * 32 + lane;
        const bool have = k_end > k_begin;
        if (have) {
            tc::mbar_wait(tmem_full, 0);
            tc::tc_fence_after_sync();
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        for (int j = 0; j < n_item_taps; ++j) {
            float* dst = p.ws + (((size_t)split * p.n_taps + tap0 + j) * p.cout + co) * p.cin + ci0;
            for (int n0 = 0; n0 < p.n_tile; n0 += 32) {
                uint32_t v[32];
                if (have) {
                    tc::tmem_ld_32x32b_x32(taddr + j * p.n_tile + n0, v);
                    tc::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) v[q] = 0u;
                }
                if (co < p.cout && ci0 + n0 < p.cin) {      // partial last blocks: zero rows / columns, never stored
#pragma unroll
                    for (int q = 0; q < 32; q += 4)
                        *reinterpret_cast<uint4*>(dst + n0 + q) = make_uint4(v[q], v[q + 1], v[q + 2], v[q + 3]);
                }
            }
        }
        tc::tc_fence_before_sync();
    }

    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after_sync();
        tc::tmem_dealloc(tmem_base, tmem_cols);
    }
}

// out[co*s_co + ci*s_ci + t*s_tap] = scale * sum_s ws[s][t][co][ci]   (fixed summation order)
__global__ void __launch_bounds__(256)
wgrad_fold_kernel(float* __restrict__ out, const float* __restrict__ ws, int splits, int n_taps, int cout, int cin,
                  long long s_co, long long s_ci, long long s_tap, float scale) {
    const long long quads = (long long)n_taps * cout * (cin / 4);
    const long long plane = (long long)n_taps * cout * cin;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < quads;
         q += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(q % (cin / 4)) * 4;
        const long long r = q / (cin / 4);
        const int co = (int)(r % cout);
        const int t = (int)(r / cout);
        const float* src = ws + ((long long)t * cout + co) * cin + ci;
        float4 acc = ld_stream_f4(reinterpret_cast<const float4*>(src));
        for (int s = 1; s < splits; ++s) {
            const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(src + s * plane));
            acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        }
        acc.x *= scale, acc.y *= scale, acc.z *= scale, acc.w *= scale;
        float* dst = out + co * s_co + ci * s_ci + t * s_tap;
        if (s_ci == 1) {
            *reinterpret_cast<float4*>(dst) = acc;
        } else {
            dst[0] = acc.x, dst[s_ci] = acc.y, dst[2 * s_ci] = acc.z, dst[3 * s_ci] = acc.w;
        }
    }
}

struct Plan {
    int tw, th, nb, tiles_x, tiles_y, tiles_b, k_tiles;
    int n_tile, cin_tiles, cout_tiles, items, splits;
    int taps_per_item, tap_groups, stages, stage_bytes;
};

int validate(const rick_wgrad_geom* g) {
    if (!g) return RICK_ERR_INVALID_ARGUMENT;
    if (g->batch < 1 || g->g_h < 1 || g->g_w < 1 || g->x_h < 1 || g->x_w < 1 || g->rows < 1 || g->cols < 1)
        return RICK_ERR_INVALID_ARGUMENT;
    if (g->n_taps < 1 || g->n_taps > 9) return RICK_ERR_INVALID_ARGUMENT;
    if (g->cout % 32 != 0 || g->cin % 32 != 0 || g->cout < 32 || g->cin < 32) return RICK_ERR_UNSUPPORTED;
    if (g->g_stride < 1 || g->g_stride > 2 || g->x_stride < 1 || g->x_stride > 2) return RICK_ERR_UNSUPPORTED;
    return RICK_OK;
}

Plan make_plan(const rick_wgrad_geom* g) {
    Plan P{};
    // 32-pixel K tile: as wide as the row allows (power of two <= 32), then rows, then samples
    int tw = 1;
    while (tw < 32 && tw < g->cols) tw <<= 1;
    int th = 1;
    while (tw * th < 32 && th < g->rows) th <<= 1;
    int nb = 32 / (tw * th);
    P.tw = tw, P.th = th, P.nb = nb;
    P.tiles_x = (int)ceil_div(g->cols, tw), P.tiles_y = (int)ceil_div(g->rows, th), P.tiles_b = (int)ceil_div(g->batch, nb);
    P.k_tiles = P.tiles_x * P.tiles_y * P.tiles_b;
    P.cin_tiles = (int)ceil_div(g->cin, kMaxN);
    P.n_tile = (int)(ceil_div(ceil_div(g->cin, P.cin_tiles), 32) * 32);     // balanced, multiple of 32
    P.cin_tiles = (int)ceil_div(g->cin, P.n_tile);
    P.cout_tiles = (int)ceil_div(g->cout, kBlockM);
    // taps that share one G tile: only when G does not move with the tap (ordinary convolutions) and Cin is one block
    bool g_fixed = true;
    for (int t = 1; t < g->n_taps; ++t) g_fixed = g_fixed && g->gy[t] == g->gy[0] && g->gx[t] == g->gx[0];
    P.taps_per_item = 1;
    if (g_fixed && P.cin_tiles == 1 && g->n_taps % 3 == 0 && 3 * P.n_tile <= kMaxCols) P.taps_per_item = 3;
    P.tap_groups = (int)ceil_div(g->n_taps, P.taps_per_item);
    P.items = P.tap_groups * P.cout_tiles * P.cin_tiles;
    P.stage_bytes = (kBlockM / 32 + P.taps_per_item * (P.n_tile / 32)) * kBlockBytes;
    P.stages = kSmemBudget / P.stage_bytes;
    if (P.stages > kMaxStages) P.stages = kMaxStages;
    // splits: ONE full wave of CTAs (a second, partial wave costs a whole extra round: one CTA per SM is resident), but
    // at least 8 K tiles (256 pixels) per CTA
    int splits = kNumSMs / P.items;
    const int by_work = P.k_tiles / 8 > 0 ? P.k_tiles / 8 : 1;
    if (splits > by_work) splits = by_work;
    if (splits < 1) splits = 1;
    if (splits > 64) splits = 64;
    P.splits = splits;
    return P;
}

}  // namespace
}  // namespace rick

extern "C" int64_t rick_conv_wgrad_workspace(const rick_wgrad_geom* g) {
    using namespace rick;
    if (validate(g) != RICK_OK) return -1;
    const Plan P = make_plan(g);
    return (int64_t)P.splits * g->n_taps * g->cout * g->cin * 4;
}

extern "C" int rick_conv_wgrad_tc(void* dw, int64_t stride_co, int64_t stride_ci, int64_t stride_tap, const void* gout,
                                  const void* x, const rick_wgrad_geom* g, void* workspace, float scale,
                                  rick_stream_t stream) {
    using namespace rick;
    const int rc = validate(g);
    if (rc != RICK_OK) return rc;
    if (!dw || !gout || !x || !workspace) return RICK_ERR_INVALID_ARGUMENT;
    if (!aligned_to(dw, 16) || !aligned_to(gout, 16) || !aligned_to(x, 16) || !aligned_to(workspace, 16))
        return RICK_ERR_ALIGNMENT;
    if (stride_ci == 1 && (stride_co % 4 != 0 || (g->n_taps > 1 && stride_tap % 4 != 0))) return RICK_ERR_ALIGNMENT;
    EncodeTiledFn encode = get_encode_tiled();
    if (!encode) return RICK_ERR_UNSUPPORTED;
    const Plan P = make_plan(g);

    WgradDev p{};
    p.batch = g->batch, p.cout = g->cout, p.cin = g->cin, p.n_taps = g->n_taps;
    p.cin_tiles = P.cin_tiles, p.n_tile = P.n_tile, p.cout_tiles = P.cout_tiles, p.items = P.items, p.splits = P.splits;
    p.taps_per_item = P.taps_per_item, p.tap_groups = P.tap_groups, p.stages = P.stages, p.stage_bytes = P.stage_bytes;
    p.tw = P.tw, p.th = P.th, p.nb = P.nb, p.tiles_x = P.tiles_x, p.tiles_y = P.tiles_y, p.tiles_b = P.tiles_b;
    p.k_tiles = P.k_tiles;
    p.g_stride = g->g_stride, p.x_stride = g->x_stride;
    for (int t = 0; t < g->n_taps; ++t) p.gy[t] = g->gy[t], p.gx[t] = g->gx[t], p.xy[t] = g->xy[t], p.xx[t] = g->xx[t];
    p.ws = static_cast<float*>(workspace);

    CUtensorMap tmap_g, tmap_x;
    // 5-D view (32 channels, W, H, B, C/32): one box = `blocks` channel blocks x 32 pixels, written block-major
    auto make_map = [&](CUtensorMap* m, const void* base, int c, int w, int h, int stride, int blocks) -> bool {
        cuuint64_t dims[5] = {32, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)g->batch, (cuuint64_t)(c / 32)};
        cuuint64_t strides[4] = {(cuuint64_t)c * 4, (cuuint64_t)c * w * 4, (cuuint64_t)c * w * h * 4, 128};
        // with a traversal stride s the box spans tw*s pixels and delivers tw of them
        cuuint32_t box[5] = {32, (cuuint32_t)(P.tw * stride), (cuuint32_t)(P.th * stride), (cuuint32_t)P.nb,
                             (cuuint32_t)blocks};
        cuuint32_t estr[5] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1, 1};
        return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    if (!make_map(&tmap_g, gout, g->cout, g->g_w, g->g_h, g->g_stride, kBlockM / 32)) return RICK_ERR_INVALID_ARGUMENT;
    if (!make_map(&tmap_x, x, g->cin, g->x_w, g->x_h, g->x_stride, P.n_tile / 32)) return RICK_ERR_INVALID_ARGUMENT;

    const size_t smem = 1024 + (size_t)P.stages * P.stage_bytes + 256;
    {
        static bool attr_done[64] = {};
        int dev = 0;
        RICK_CUDA_TRY(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {
            RICK_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               1024 + kSmemBudget + 256));
            if (dev >= 0 && dev < 64) attr_done[dev] = true;
        }
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    conv_wgrad_kernel<<<P.items * P.splits, kThreads, smem, st>>>(tmap_g, tmap_x, p);
    RICK_CHECK_LAUNCH();
    const long long quads = (long long)g->n_taps * g->cout * (g->cin / 4);
    long long blocks = ceil_div(quads, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    wgrad_fold_kernel<<<(unsigned)blocks, 256, 0, st>>>(static_cast<float*>(dw), p.ws, P.splits, g->n_taps, g->cout,
                                                        g->cin, stride_co, stride_ci, stride_tap,
                                                        scale * tc::kTf32TruncationComp);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}
