// Diagnostics for the tcgen05 path (not used by the product): a single-CTA, single-k-block UMMA whose B operand
// descriptor starts `shift` rows into a 128-byte-swizzled tile.  It answers, on real hardware, whether a K-major
// SWIZZLE_128B operand may start at a row that is not a multiple of 8 (needed to reuse one shared-memory halo tile
// for all 3x3 filter taps) and which `base_offset` encoding that requires.
#include "common.cuh"
#include "tc_common.cuh"

namespace rick {
namespace {

constexpr int kM = 128, kN = 64, kK = 32, kBRows = 96;

__global__ void __launch_bounds__(128, 1)
umma_shift_probe(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, float* out,
                 int shift, int base_offset_mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_tile = smem;                       // 128 rows x 128 B
    uint8_t* b_tile = smem + kM * 128;            // 96 rows x 128 B
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(b_tile + kBRows * 128);
    uint64_t* bar_mma = bar_load + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tc::mbar_init(bar_load, 1);
        tc::mbar_init(bar_mma, 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) {
        tc::tmem_alloc(tmem_slot, 64);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0) {
        tc::mbar_arrive_expect_tx(bar_load, (kM + kBRows) * 128);
        tc::tma_load_3d(a_tile, &tmap_a, bar_load, 0, 0, 0);
        tc::tma_load_3d(b_tile, &tmap_b, bar_load, 0, 0, 0);
        tc::mbar_wait(bar_load, 0);
        tc::tc_fence_after_sync();
        const uint32_t b_addr = tc::smem_u32(b_tile) + shift * 128;
        const uint32_t bo = base_offset_mode ? ((b_addr >> 7) & 7) : 0;
        const uint64_t a_desc = tc::umma_desc_k_sw128(tc::smem_u32(a_tile));
        const uint64_t b_desc = tc::umma_desc_k_sw128(b_addr, bo);
        const uint32_t idesc = tc::umma_idesc_tf32(kM, kN);
        for (int k = 0; k < kK / 8; ++k) tc::umma_tf32_ss(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0);
        tc::umma_commit(bar_mma);
    }
    __syncwarp();
    tc::mbar_wait(bar_mma, 0);
    tc::tc_fence_after_sync();
    for (int n0 = 0; n0 < kN; n0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + n0, v);
        tc::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * kN + n0 + j] = __uint_as_float(v[j]);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        tc::tc_fence_after_sync();
        tc::tmem_dealloc(tmem, 64);
    }
}

}  // namespace
}  // namespace rick

// a: (128, 32) fp32, b: (96, 32) fp32, out: (128, 64) fp32 = a @ b[shift : shift + 64].T (TF32)
extern "C" int rick_debug_umma_shift(float* out, const float* a, const float* b, int shift, int base_offset_mode,
                                     rick_stream_t stream) {
    using namespace rick;
    if (!out || !a || !b || shift < 0 || shift + kN > kBRows) return RICK_ERR_INVALID_ARGUMENT;
    EncodeTiledFn encode = get_encode_tiled();
    if (!encode) return RICK_ERR_UNSUPPORTED;
    CUtensorMap ta, tb;
    cuuint32_t estr[3] = {1, 1, 1};
    {
        cuuint64_t dims[3] = {kK, kM, 1};
        cuuint64_t strides[2] = {kK * 4, (cuuint64_t)kK * kM * 4};
        cuuint32_t box[3] = {kK, kM, 1};
        if (encode(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return RICK_ERR_INVALID_ARGUMENT;
    }
    {
        cuuint64_t dims[3] = {kK, kBRows, 1};
        cuuint64_t strides[2] = {kK * 4, (cuuint64_t)kK * kBRows * 4};
        cuuint32_t box[3] = {kK, kBRows, 1};
        if (encode(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(b), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return RICK_ERR_INVALID_ARGUMENT;
    }
    const size_t smem = 1024 + (kM + kBRows) * 128 + 64;
    RICK_CUDA_TRY(cudaFuncSetAttribute(umma_shift_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_shift_probe<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(ta, tb, out, shift, base_offset_mode);
    RICK_CHECK_LAUNCH();
    return RICK_OK;
}
