// Hardware probe (diagnostics, not part of the product library): tcgen05.mma.kind::tf32 with MN-major shared-memory
// operands, as the weight-gradient / data-gradient kernels need them (activations NHWC => the GEMM-K axis "pixels" is
// the strided one).  One CTA, one accumulator:  D[m][n] = sum_k A[k][m] * B[k][n]  with
//     A in global memory as [K][M] (M contiguous), B as [K][N] (N contiguous)           (MN-major)
//  or A as [M][K], B as [N][K] (K contiguous)                                            (K-major, the proven path)
// MN-major tiles are staged by TMA as blocks of 32 contiguous elements (128 B, SWIZZLE_128B) x K rows, block j at
// tile + j * K * 128 B, which is the canonical "((8,n),(8,k)):((1,LBO),(8,SBO))" layout (in 16-byte units) CUTLASS
// documents for Major::MN / SWIZZLE_128B.  The probe tries the LBO / SBO assignments and prints the error of each.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I rick_b200/csrc -o scripts/build/probe_mnmajor scripts/probe_mnmajor.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>

#include "tc_common.cuh"

using namespace rick;

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
}  // namespace rick

constexpr int kM = 128, kN = 64, kK = 64;   // K = 64: 8 MMA k-steps, two 1024 B atoms per 16 rows

struct ProbeCfg {
    int a_mn, b_mn;          // 1 = MN-major operand
    unsigned lbo, sbo;       // bytes, MN-major descriptors
    unsigned kstep_bytes;    // descriptor start-address advance per k-step of 8 (MN-major)
    unsigned layout_type;    // UMMA layout type of the MN-major descriptors: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3ffff) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout_type) << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, float* out,
             const ProbeCfg cfg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_tile = smem;                          // kM * kK * 4 bytes
    uint8_t* b_tile = smem + kM * kK * 4;            // kN * kK * 4 bytes
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(b_tile + kN * kK * 4);
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
        tc::mbar_arrive_expect_tx(bar_load, (kM + kN) * kK * 4);
        if (cfg.a_mn) {
            for (int j = 0; j < kM / 32; ++j)        // block j: 32 m-columns x kK rows
                tc::tma_load_3d(a_tile + j * kK * 128, &tmap_a, bar_load, j * 32, 0, 0);
        } else {
            for (int j = 0; j < kK / 32; ++j)        // k-block j: kM rows x 32 k
                tc::tma_load_3d(a_tile + j * kM * 128, &tmap_a, bar_load, j * 32, 0, 0);
        }
        if (cfg.b_mn) {
            for (int j = 0; j < kN / 32; ++j)
                tc::tma_load_3d(b_tile + j * kK * 128, &tmap_b, bar_load, j * 32, 0, 0);
        } else {
            for (int j = 0; j < kK / 32; ++j)
                tc::tma_load_3d(b_tile + j * kN * 128, &tmap_b, bar_load, j * 32, 0, 0);
        }
        tc::mbar_wait(bar_load, 0);
        tc::tc_fence_after_sync();
        const uint32_t idesc = tc::umma_idesc_tf32(kM, kN) | (cfg.a_mn ? (1u << 15) : 0u) | (cfg.b_mn ? (1u << 16) : 0u);
        for (int k = 0; k < kK / 8; ++k) {
            uint64_t a_desc, b_desc;
            if (cfg.a_mn)
                a_desc = make_desc(tc::smem_u32(a_tile) + k * cfg.kstep_bytes, cfg.lbo, cfg.sbo, cfg.layout_type);
            else   // K-major: k-block (k / 4) of 32 floats, 32 B per k-step inside the swizzled row
                a_desc = tc::umma_desc_k_sw128(tc::smem_u32(a_tile) + (k / 4) * kM * 128) + 2 * (k % 4);
            if (cfg.b_mn)
                b_desc = make_desc(tc::smem_u32(b_tile) + k * cfg.kstep_bytes, cfg.lbo, cfg.sbo, cfg.layout_type);
            else
                b_desc = tc::umma_desc_k_sw128(tc::smem_u32(b_tile) + (k / 4) * kN * 128) + 2 * (k % 4);
            tc::umma_tf32_ss(tmem, a_desc, b_desc, idesc, k != 0);
        }
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

static float tf32_round(float x) {   // round-to-nearest-even on the 13 dropped bits is not what the MMA does (it
    uint32_t u;                       // truncates); accept either by comparing with a loose tolerance
    memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}

static bool encode2d(CUtensorMap* m, const float* base, int inner, int outer, int box_inner, int box_outer,
                     CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn encode = get_encode_tiled();
    if (!encode) return false;
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, 1};
    cuuint64_t strides[2] = {(cuuint64_t)inner * 4, (cuuint64_t)inner * outer * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int main() {
    std::vector<float> A(kK * kM), B(kK * kN);          // logical A[k][m], B[k][n]
    srand(1);
    for (auto& v : A) v = tf32_round((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : B) v = tf32_round((rand() % 2001 - 1000) / 1000.f);
    std::vector<float> want(kM * kN, 0.f);
    for (int m = 0; m < kM; ++m)
        for (int n = 0; n < kN; ++n) {
            double s = 0;
            for (int k = 0; k < kK; ++k) s += (double)A[k * kM + m] * B[k * kN + n];
            want[m * kN + n] = (float)s;
        }
    // device copies in both layouts
    std::vector<float> At(kM * kK), Bt(kN * kK);        // K-major: At[m][k], Bt[n][k]
    for (int k = 0; k < kK; ++k) {
        for (int m = 0; m < kM; ++m) At[m * kK + k] = A[k * kM + m];
        for (int n = 0; n < kN; ++n) Bt[n * kK + k] = B[k * kN + n];
    }
    float *dA, *dB, *dAt, *dBt, *dOut;
    cudaMalloc(&dA, A.size() * 4), cudaMalloc(&dB, B.size() * 4), cudaMalloc(&dAt, At.size() * 4);
    cudaMalloc(&dBt, Bt.size() * 4), cudaMalloc(&dOut, kM * kN * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dAt, At.data(), At.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dBt, Bt.data(), Bt.size() * 4, cudaMemcpyHostToDevice);

    CUtensorMap ta_mn, tb_mn, ta_mn32, tb_mn32, ta_k, tb_k;
    bool ok = encode2d(&ta_mn, dA, kM, kK, 32, kK) && encode2d(&tb_mn, dB, kN, kK, 32, kK) &&
              encode2d(&ta_mn32, dA, kM, kK, 32, kK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) &&
              encode2d(&tb_mn32, dB, kN, kK, 32, kK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) &&
              encode2d(&ta_k, dAt, kK, kM, 32, kM) && encode2d(&tb_k, dBt, kK, kN, 32, kN);
    if (!ok) { printf("tensor map encode failed\n"); return 2; }
    const size_t smem = 1024 + (kM + kN) * kK * 4 + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);

    struct Variant { const char* name; unsigned lbo, sbo, kstep, layout; };
    const Variant variants[] = {
        {"SW128 lbo=block(K*128) sbo=1024", kK * 128, 1024, 1024, 2},
        {"SW128_BASE32B lbo=block(K*128) sbo=512", kK * 128, 512, 1024, 1},
        {"SW128_BASE32B lbo=512 sbo=block", 512, kK * 128, 1024, 1},
        {"SW128_BASE32B lbo=block sbo=1024", kK * 128, 1024, 1024, 1},
        {"SW128_BASE32B lbo=block sbo=block", kK * 128, kK * 128, 1024, 1},
    };
    const int majors[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};
    std::vector<float> got(kM * kN);
    for (auto& mj : majors) {
        for (auto& v : variants) {
            if (!mj[0] && !mj[1] && &v != &variants[0]) continue;     // K-major baseline once
            ProbeCfg cfg{mj[0], mj[1], v.lbo, v.sbo, v.kstep, v.layout};
            const bool b32 = v.layout == 1;
            cudaMemset(dOut, 0, kM * kN * 4);
            probe_kernel<<<1, 128, smem>>>(mj[0] ? (b32 ? ta_mn32 : ta_mn) : ta_k, mj[1] ? (b32 ? tb_mn32 : tb_mn) : tb_k, dOut, cfg);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("A %s B %s  [%s]: CUDA error %s\n", mj[0] ? "MN" : "K ", mj[1] ? "MN" : "K ", v.name,
                       cudaGetErrorString(e));
                return 3;     // context is gone after a trap
            }
            cudaMemcpy(got.data(), dOut, kM * kN * 4, cudaMemcpyDeviceToHost);
            double worst = 0, ref = 0;
            int bad = 0;
            for (int i = 0; i < kM * kN; ++i) {
                worst = fmax(worst, fabs((double)got[i] - want[i]));
                ref = fmax(ref, fabs((double)want[i]));
                if (fabs((double)got[i] - want[i]) > 1e-3 * 8) ++bad;
            }
            printf("A %s B %s  [%s]: max abs err %.3e (max |want| %.2f), %d / %d wrong  %s   got[0..3] %.3f %.3f %.3f %.3f want %.3f %.3f %.3f %.3f\n", mj[0] ? "MN" : "K ",
                   mj[1] ? "MN" : "K ", v.name, worst, ref, bad, kM * kN, bad == 0 ? "OK" : "--", got[0], got[1], got[2], got[3], want[0], want[1], want[2], want[3]);
        }
    }
    return 0;
}
