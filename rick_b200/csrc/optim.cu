// Fused multi-tensor optimiser step for the adaptation loop (sm_100a): filter masks + Adam + EMA in ONE pass over
// every parameter of a network, one launch per network.
//
// Replaces, per optimiser step of the reference loop (train_dynamic_update_prune.py):
//   :427-437 / :521-539   grad[freeze] = 0; param[zero] = 0; grad[zero] = 0   (~200-400 index_put launches)
//   torch.optim.Adam.step  (betas = (0, 0.99**ratio), train:916-925)
//   :68-73, 697-698       accumulate(g_ema, g_module, 0.5 ** (32 / 10000))     (2 launches per parameter)
// Each element is read and written once: p, g, m, v (+ ema) in, p, m, v (+ ema) out -- 28 B (36 B with EMA) per
// parameter; HBM-bound.  Masks are per-filter bytes (row = filter = outermost dimension of the tensor), so a row
// lookup is one integer divide per 128-bit access.  Gradients are NOT written back: nothing reads them after the step
// (the loop clears them before the next backward), so the masked gradient only exists in registers.
#include "common.cuh"

namespace rick {
namespace {

constexpr int kMaxAdamTensors = 128;
constexpr int kAdamChunk = 256 * 8;   // elements per CTA

struct AdamEntry {
    float* p;
    const float* g;
    float* m;
    float* v;
    float* ema;
    const uint8_t* state;
    const uint8_t* zero;
    const float* step;    // this tensor's own step count (torch.optim.Adam keeps one per parameter)
    long long total;
    unsigned inner;
};
struct AdamTable {
    AdamEntry e[kMaxAdamTensors];
    int block_end[kMaxAdamTensors];
    int count;
    float lr, beta1, beta2, eps, decay;
};

struct AdamCoef {
    float beta1, beta2, one_m_beta1, one_m_beta2, step_size, inv_bc2_sqrt, eps, decay, one_m_decay;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float& ema, bool has_g, bool has_ema,
                                          bool z, bool f, const AdamCoef& c) {
    if (z) p = 0.f;                                       // pruned filter: weight pinned to zero (train:431, 530)
    if (has_g) {
        if (z || f) g = 0.f;                              // frozen / pruned filter: no gradient (train:427, 433)
        m = c.beta1 * m + c.one_m_beta1 * g;
        v = c.beta2 * v + c.one_m_beta2 * g * g;
        const float denom = sqrtf(v) * c.inv_bc2_sqrt + c.eps;
        p -= (c.step_size * m) / denom;
    }
    if (has_ema) ema = fmaf(p, c.one_m_decay, ema * c.decay);   // accumulate(): ema * decay + (1 - decay) * p
}

__global__ void __launch_bounds__(256) adam_mask_ema_kernel(const __grid_constant__ AdamTable tab) {
    int t = 0;
    while (t < tab.count - 1 && (int)blockIdx.x >= tab.block_end[t]) ++t;
    const int local_block = blockIdx.x - (t ? tab.block_end[t - 1] : 0);
    const AdamEntry& e = tab.e[t];
    const float step = e.step ? __ldg(e.step) : 1.f;
    AdamCoef c;
    c.beta1 = tab.beta1, c.beta2 = tab.beta2, c.one_m_beta1 = 1.f - tab.beta1, c.one_m_beta2 = 1.f - tab.beta2;
    const float bc1 = 1.f - (tab.beta1 == 0.f ? 0.f : powf(tab.beta1, step));
    const float bc2 = 1.f - powf(tab.beta2, step);
    c.step_size = tab.lr / bc1, c.inv_bc2_sqrt = 1.f / sqrtf(bc2), c.eps = tab.eps;
    c.decay = tab.decay, c.one_m_decay = 1.f - tab.decay;

    const bool has_g = e.g != nullptr, has_ema = e.ema != nullptr;
    const long long begin = (long long)local_block * kAdamChunk;
    const long long end = min(begin + (long long)kAdamChunk, e.total);
    const uintptr_t align = reinterpret_cast<uintptr_t>(e.p) | reinterpret_cast<uintptr_t>(e.g) |
                            reinterpret_cast<uintptr_t>(e.m) | reinterpret_cast<uintptr_t>(e.v) |
                            reinterpret_cast<uintptr_t>(e.ema);
    const bool masked = e.state || e.zero;
    if ((align & 15) == 0 && (e.inner & 3) == 0 && (e.total & 3) == 0) {
        // 128-bit path: a float4 never straddles two filters because inner % 4 == 0
        for (long long i = begin + threadIdx.x * 4; i < end; i += 256 * 4) {
            bool z = false, f = false;
            if (masked) {
                const long long row = i / e.inner;
                z = e.zero && e.zero[row];
                f = e.state && (e.state[row] & 1);
            }
            float4 p = *reinterpret_cast<const float4*>(e.p + i);
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f), m = g, v = g, a = g;
            if (has_g) {
                g = ld_stream_f4(reinterpret_cast<const float4*>(e.g + i));
                m = *reinterpret_cast<const float4*>(e.m + i);
                v = *reinterpret_cast<const float4*>(e.v + i);
            }
            if (has_ema) a = *reinterpret_cast<const float4*>(e.ema + i);
            adam_elem(p.x, g.x, m.x, v.x, a.x, has_g, has_ema, z, f, c);
            adam_elem(p.y, g.y, m.y, v.y, a.y, has_g, has_ema, z, f, c);
            adam_elem(p.z, g.z, m.z, v.z, a.z, has_g, has_ema, z, f, c);
            adam_elem(p.w, g.w, m.w, v.w, a.w, has_g, has_ema, z, f, c);
            if (has_g || z) *reinterpret_cast<float4*>(e.p + i) = p;
            if (has_g) {
                *reinterpret_cast<float4*>(e.m + i) = m;
                *reinterpret_cast<float4*>(e.v + i) = v;
            }
            if (has_ema) *reinterpret_cast<float4*>(e.ema + i) = a;
        }
    } else {
        for (long long i = begin + threadIdx.x; i < end; i += 256) {
            bool z = false, f = false;
            if (masked) {
                const long long row = i / e.inner;
                z = e.zero && e.zero[row];
                f = e.state && (e.state[row] & 1);
            }
            float p = e.p[i], g = 0.f, m = 0.f, v = 0.f, a = 0.f;
            if (has_g) g = e.g[i], m = e.m[i], v = e.v[i];
            if (has_ema) a = e.ema[i];
            adam_elem(p, g, m, v, a, has_g, has_ema, z, f, c);
            if (has_g || z) e.p[i] = p;
            if (has_g) e.m[i] = m, e.v[i] = v;
            if (has_ema) e.ema[i] = a;
        }
    }
}

// ---- out[t] = in[t] * scale[t] for a whole network's weights in one launch --------------------------------------
// The equalised-lr multipliers (EqualConv2d / EqualLinear: weight * 1/sqrt(fan_in), bias * lr_mul,
// model_probe_tune.py:110-124, 155-165) are ~45 separate 2-microsecond element-wise launches per network pass and as
// many again in backward; as a multi-tensor kernel they are one launch each way.
constexpr int kMaxScaleTensors = 160;
constexpr int kScaleChunk = 256 * 16;
struct ScaleTable {
    float* out[kMaxScaleTensors];
    const float* in[kMaxScaleTensors];
    long long numel[kMaxScaleTensors];
    float scale[kMaxScaleTensors];
    int block_end[kMaxScaleTensors];
    int count;
};

__global__ void __launch_bounds__(256) scale_multi_kernel(const __grid_constant__ ScaleTable tab) {
    int t = 0;
    while (t < tab.count - 1 && (int)blockIdx.x >= tab.block_end[t]) ++t;
    const int local_block = blockIdx.x - (t ? tab.block_end[t - 1] : 0);
    float* __restrict__ out = tab.out[t];
    const float* __restrict__ in = tab.in[t];
    const float s = tab.scale[t];
    const long long n = tab.numel[t];
    const long long begin = (long long)local_block * kScaleChunk;
    const long long end = min(begin + (long long)kScaleChunk, n);
    if (((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(in)) & 15) == 0 && (n & 3) == 0) {
        for (long long i = begin + threadIdx.x * 4; i < end; i += 256 * 4) {
            float4 v = *reinterpret_cast<const float4*>(in + i);
            v.x *= s, v.y *= s, v.z *= s, v.w *= s;
            *reinterpret_cast<float4*>(out + i) = v;
        }
    } else {
        for (long long i = begin + threadIdx.x; i < end; i += 256) out[i] = in[i] * s;
    }
}

}  // namespace
}  // namespace rick

extern "C" int rick_scale_multi(float* const* out, const float* const* in, const float* scale, const int64_t* numel,
                                int count, rick_stream_t stream) {
    using namespace rick;
    if (!out || !in || !scale || !numel || count < 0) return RICK_ERR_INVALID_ARGUMENT;
    for (int base = 0; base < count; base += kMaxScaleTensors) {
        ScaleTable tab{};
        const int n = (count - base < kMaxScaleTensors) ? count - base : kMaxScaleTensors;
        long long blocks = 0;
        int used = 0;
        for (int i = 0; i < n; ++i) {
            const int k = base + i;
            if (numel[k] < 0 || !out[k] || !in[k]) return RICK_ERR_INVALID_ARGUMENT;
            if (numel[k] == 0) continue;
            tab.out[used] = out[k], tab.in[used] = in[k], tab.numel[used] = numel[k], tab.scale[used] = scale[k];
            blocks += ceil_div(numel[k], kScaleChunk);
            if (blocks > 0x7fffffffLL) return RICK_ERR_OVERFLOW;
            tab.block_end[used] = (int)blocks;
            ++used;
        }
        if (!used) continue;
        tab.count = used;
        scale_multi_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(tab);
        RICK_CHECK_LAUNCH();
    }
    return RICK_OK;
}

extern "C" int rick_adam_mask_ema(float* const* param, const float* const* grad, float* const* exp_avg,
                                  float* const* exp_avg_sq, float* const* ema, const uint8_t* const* state,
                                  const uint8_t* const* zero, const float* const* step, const int64_t* rows,
                                  const int64_t* inner, int count, float lr, float beta1, float beta2, float eps, float ema_decay,
                                  rick_stream_t stream) {
    using namespace rick;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !ema || !state || !zero || !rows || !inner || !step || count < 0)
        return RICK_ERR_INVALID_ARGUMENT;
    if (!(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f)) return RICK_ERR_INVALID_ARGUMENT;
    for (int base = 0; base < count; base += kMaxAdamTensors) {
        AdamTable tab{};
        const int n = (count - base < kMaxAdamTensors) ? count - base : kMaxAdamTensors;
        long long blocks = 0;
        int used = 0;
        for (int i = 0; i < n; ++i) {
            const int k = base + i;
            if (rows[k] < 0 || inner[k] < 1 || inner[k] > 0xffffffffLL) return RICK_ERR_INVALID_ARGUMENT;
            if (!param[k]) return RICK_ERR_INVALID_ARGUMENT;
            if (grad[k] && (!exp_avg[k] || !exp_avg_sq[k] || !step[k])) return RICK_ERR_INVALID_ARGUMENT;
            if (rows[k] == 0 || (!grad[k] && !ema[k] && !zero[k])) continue;
            AdamEntry& e = tab.e[used];
            e.p = param[k], e.g = grad[k], e.m = exp_avg[k], e.v = exp_avg_sq[k], e.ema = ema[k];
            e.state = state[k], e.zero = zero[k], e.step = step[k];
            e.total = rows[k] * inner[k], e.inner = (unsigned)inner[k];
            blocks += ceil_div(e.total, kAdamChunk);
            if (blocks > 0x7fffffffLL) return RICK_ERR_OVERFLOW;
            tab.block_end[used] = (int)blocks;
            ++used;
        }
        if (!used) continue;
        tab.count = used;
        tab.lr = lr, tab.beta1 = beta1, tab.beta2 = beta2, tab.eps = eps, tab.decay = ema_decay;
        adam_mask_ema_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(tab);
        RICK_CHECK_LAUNCH();
    }
    return RICK_OK;
}
