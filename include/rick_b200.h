/* rick_b200 -- C ABI of the B200-native RICK / StyleGAN2 hot path.
 *
 * Drop-in boundary.  Every entry point takes raw device pointers,
 * explicit sizes, a dtype enum and a cudaStream_t (passed as void*), launches asynchronously
 * on that stream, never allocates, never synchronises, and returns an int status
 * (0 = RICK_OK).  No exceptions and no torch types cross this boundary.  All functions
 * are re-entrant and keep no global mutable state; the device is the calling thread's
 * current CUDA device (the Python shim sets it from the tensor).
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   rick_upfirdn2d            op/upfirdn2d.cpp:12-19  upfirdn2d(input, kernel, up_x, up_y, down_x, down_y,
 *                             pad_x0, pad_x1, pad_y0, pad_y1) -> op/upfirdn2d_kernel.cu:209-369
 *   rick_bias_act             op/fused_bias_act.cpp:11-17  fused_bias_act(input, bias, refer, act, grad, alpha, scale)
 *                             -> op/fused_bias_act_kernel.cu:52-98
 *   rick_bias_act_bwd         op/fused_act.py:19-39  (kernel mode act=3, grad=1) + the separate
 *                             ``grad_input.sum(dim)`` bias reduction, fused into one pass
 *   rick_fisher_*             train_dynamic_update_prune.py:252-269 (grad**2 accumulation / averaging on host NumPy)
 *   rick_filter_fim           train:282, 291-293, 339-341, 349 (per-filter ``ndarray.mean`` [+ bias] / 2)
 *   rick_percentile           train:285-286, 298-299, 352-353 (np.percentile, 'linear')
 *   rick_decide               train:312-314, 324-330, 368-384 (np.where index sets) + 386-393 (cumulative prune union)
 *   rick_mask_apply           train:427-437, 482-492, 521-539, 566-585 (index_put on param / grad per filter)
 *   rick_adam_mask_ema        the same masks + torch.optim.Adam.step (train:916-925) + accumulate() (train:68-73, 697-698)
 *   rick_scale_multi          the equalised-lr multipliers ``weight * scale`` / ``bias * lr_mul`` of every EqualConv2d /
 *                             EqualLinear (gan_training/models/model_probe_tune.py:124, 160-164), one launch per network pass
 *   rick_conv_tc              gan_training/models/model_probe_tune.py:243-284 (ModulatedConv2d: modulate, demodulate,
 *                             grouped conv / transposed conv) and :122-128 (EqualConv2d) -- the forward AND data-gradient
 *                             convolutions as one tcgen05 implicit-GEMM kernel, rick_b200/csrc/conv_tc.cu
 *   rick_conv_wgrad_tc        the weight gradient of the same convolutions (autograd's cuDNN wgrad in the reference),
 *                             rick_b200/csrc/conv_wgrad.cu
 */
#ifndef RICK_B200_H
#define RICK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RICK_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define RICK_API __attribute__((visibility("default")))
#else
#define RICK_API
#endif

typedef void* rick_stream_t; /* cudaStream_t */

enum rick_status {
    RICK_OK = 0,
    RICK_ERR_INVALID_ARGUMENT = 1, /* null pointer, non-positive size, bad enum */
    RICK_ERR_UNSUPPORTED = 2,      /* valid request this build has no kernel for */
    RICK_ERR_OVERFLOW = 3,         /* a size does not fit the kernel's index type */
    RICK_ERR_CUDA = 4,             /* launch / runtime error, see rick_last_cuda_error() */
    RICK_ERR_ALIGNMENT = 5         /* pointer not aligned for its element type */
};

enum rick_dtype { RICK_F32 = 0, RICK_BF16 = 1 };

/* bias-act activation / derivative selectors, numbered as the reference kernel's (act, grad) pair */
enum rick_act { RICK_ACT_LINEAR = 1, RICK_ACT_LRELU = 3 };

RICK_API int rick_abi_version(void);
RICK_API const char* rick_status_string(int status);
/* cudaGetLastError()-style text of the most recent CUDA failure seen by the calling thread ("" if none) */
RICK_API const char* rick_last_cuda_error(void);
/* number of kernels launched by this library in this process so far (statistics for bench.py; monotonic) */
RICK_API unsigned long long rick_launch_count(void);

/* ------------------------------------------------------------------------------------------- upfirdn2d
 * in : (major, in_h, in_w, minor) contiguous     out: (major, out_h, out_w, minor) contiguous
 * taps: (kh, kw) float32 on the device.  flip_taps != 0 uses taps[kh-1-i][kw-1-j] instead (what the
 * reference obtains with torch.flip for the backward pass, op/upfirdn2d.py:101).
 * out_h = (in_h*up_y + pad_y0 + pad_y1 - kh) / down_y + 1 (floor), same for w.
 * minor == 1 is NCHW planes (major = N*C); minor == C is channels-last. */
RICK_API int rick_upfirdn2d_out_size(int in_size, int k, int up, int down, int pad0, int pad1);
RICK_API int rick_upfirdn2d(void* out, const void* in, const float* taps, int64_t major, int in_h, int in_w, int minor,
                   int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                   int pad_y1, int flip_taps, int dtype, rick_stream_t stream);

/* ------------------------------------------------------------------------------------------- bias-act
 * out[i] = f(x[i] + bias[(i / step_b) % size_b]) * scale with
 *   act=LINEAR: f(v)=v;  act=LRELU, grad=0: v>0 ? v : v*alpha;  grad=1: ref[i]>0 ? v : v*alpha;  grad=2: 0.
 * bias == NULL / ref == NULL mean "absent" (the reference passes empty tensors). */
RICK_API int rick_bias_act(void* out, const void* x, const void* bias, const void* ref, int64_t n, int64_t step_b,
                  int64_t size_b, int act, int grad, float alpha, float scale, int dtype, rick_stream_t stream);

/* Fused backward of fused_leaky_relu: grad_in = (out>0 ? g : g*alpha)*scale and, in the same pass,
 * grad_bias[c] = sum over (n, hw) of grad_in, deterministically (two-stage, no float atomics).
 * Tensors are (N, C, HW) contiguous; grad_bias is float32 (C); workspace must hold
 * rick_bias_act_bwd_workspace(N, C, HW) bytes. */
RICK_API int64_t rick_bias_act_bwd_workspace(int64_t n, int64_t c, int64_t hw);
RICK_API int rick_bias_act_bwd(void* grad_in, float* grad_bias, void* workspace, const void* grad_out, const void* out_saved,
                      int64_t n, int64_t c, int64_t hw, float alpha, float scale, int dtype, rick_stream_t stream);

/* Channels-last variant: tensors are (rows = N*H*W, C) contiguous, the bias channel is the innermost index.
 * workspace: rick_bias_act_bwd_nhwc_workspace(rows, c) bytes.  fp32 only. */
RICK_API int64_t rick_bias_act_bwd_nhwc_workspace(int64_t rows, int64_t c);
RICK_API int rick_bias_act_bwd_nhwc(void* grad_in, float* grad_bias, void* workspace, const void* grad_out,
                                    const void* out_saved, int64_t rows, int64_t c, float alpha, float scale,
                                    rick_stream_t stream);

/* ------------------------------------------------------------------------------------------- Fisher
 * acc/grad pointer tables are HOST arrays of DEVICE pointers (count entries); they are copied into kernel
 * parameters, so no device-side table and no H2D copy is needed.
 * first != 0:  acc = grad*grad       else:  acc = acc + grad*grad   (float32, product rounded before the add,
 * exactly like ``fisher += (g ** 2).cpu().numpy()``). */
RICK_API int rick_fisher_accum(float* const* acc, const float* const* grad, const int64_t* numel, int count, int first,
                      rick_stream_t stream);
/* acc[i] = acc[i] / divisor  (float32 division, ``fisher /= num_fisher_img * batch``) */
RICK_API int rick_fisher_divide(float* const* acc, const int64_t* numel, int count, float divisor, rick_stream_t stream);

/* Per-filter FIM of one layer: rows = filters, len = elements per filter (contiguous).
 * fim[r] = mean_f32(fisher_w[r, :])                          if fisher_b == NULL
 *        = (mean_f32(fisher_w[r, :]) + fisher_b[r]) / 2      otherwise
 * mean_f32 reproduces NumPy's float32 pairwise summation order bit for bit. */
RICK_API int rick_filter_fim(float* fim, const float* fisher_w, const float* fisher_b, int64_t rows, int64_t len,
                    rick_stream_t stream);

/* np.percentile(float64(fim[0:n]), q[j]) for j < nq (nq <= 8), method 'linear', by radix-select of the
 * two neighbouring order statistics; q is a HOST array; lines is a DEVICE array of nq doubles. */
RICK_API int rick_percentile(double* lines, const float* fim, int64_t n, const double* q, int nq, rick_stream_t stream);

/* state[i] = (fim>cut ? 1:0) | (prune ? 2:0) | (fine-tune ? 4:0) with cut = lines[0], pruneline = lines[1].
 * flags bit 0 (RICK_DECIDE_CLOSED_LOW) selects the reference's D-skip comparisons (train:382-384); bit 1
 * (RICK_DECIDE_COMPARE_F32) compares in float32 against the thresholds rounded to float32 -- NumPy < 2 value-based
 * casting, what the reference's pinned NumPy 1.23.1 does -- instead of in float64 (NumPy >= 2 promotion).  If
 * zero_mask != NULL it is OR-ed with the prune bit (cumulative prune set, train:386-393); pass reset_zero != 0 on
 * the first round. */
#define RICK_DECIDE_CLOSED_LOW 1
#define RICK_DECIDE_COMPARE_F32 2
RICK_API int rick_decide(uint8_t* state, uint8_t* zero_mask, const float* fim, int64_t n, const double* lines,
                int flags, int reset_zero, rick_stream_t stream);

/* One launch for a whole model.  Entry t: param/grad are (rows[t], inner[t]) contiguous float32, state / zero are
 * per-row bytes (NULL = no such set for this tensor).  grad rows with (state&1) or zero are cleared, param rows
 * with zero are cleared.  Tables are HOST arrays. */
RICK_API int rick_mask_apply(float* const* param, float* const* grad, const uint8_t* const* state,
                    const uint8_t* const* zero, const int64_t* rows, const int64_t* inner, int count,
                    rick_stream_t stream);

/* Fused multi-tensor optimiser step: filter masks + Adam + EMA in one pass, one launch per network.  Replaces, per
 * optimiser step of the reference loop, the index_put mask application (train:427-437, 482-492, 521-539, 566-585),
 * torch.optim.Adam.step (betas (0, 0.99**ratio), train:916-925) and, when ema[t] is given, accumulate() (train:68-73,
 * 697-698).  HOST tables of length count; entry t describes rows[t] x inner[t] contiguous float32 elements:
 *   param[t]                 updated in place
 *   grad[t]                  NULL = no optimiser update for this tensor (EMA / pruning only); never written
 *   exp_avg[t], exp_avg_sq[t] Adam moments (required when grad[t] is given)
 *   ema[t]                   NULL = no EMA; else ema = ema * ema_decay + (1 - ema_decay) * param (after the update)
 *   state[t], zero[t]        per-row bytes as for rick_mask_apply (NULL = unmasked): rows with (state & 1) or zero take
 *                            a zero gradient, rows with zero have param set to 0
 *   step[t]                  DEVICE pointer to this tensor's float step count t >= 1, already incremented by the caller
 *                            (torch.optim.Adam keeps one count per parameter: a tensor that starts receiving
 *                            gradients late starts its bias correction late; device-resident so that the launch can
 *                            be recorded into a CUDA graph).  Required when grad[t] is given. */
RICK_API int rick_adam_mask_ema(float* const* param, const float* const* grad, float* const* exp_avg,
                       float* const* exp_avg_sq, float* const* ema, const uint8_t* const* state,
                       const uint8_t* const* zero, const float* const* step, const int64_t* rows, const int64_t* inner,
                       int count, float lr, float beta1, float beta2, float eps, float ema_decay,
                       rick_stream_t stream);

/* out[t] = in[t] * scale[t] over count dense float32 tensors in one launch: the equalised-lr multipliers of EqualConv2d /
 * EqualLinear (model_probe_tune.py:124, 160-164: ``self.weight * self.scale``, ``self.bias * self.lr_mul``) for a whole
 * network pass, and their backward (grad * scale).  Tables are HOST arrays. */
RICK_API int rick_scale_multi(float* const* out, const float* const* in, const float* scale, const int64_t* numel, int count,
                     rick_stream_t stream);

/* ------------------------------------------------------------------------------------------- tcgen05 convolution
 * Implicit-GEMM convolution on NHWC fp32 activations (TF32 tensor-core math, fp32 accumulate), replacing the grouped
 * cuDNN convolutions of ModulatedConv2d (gan_training/models/model_probe_tune.py:265, 274, 280) and the dense ones
 * of EqualConv2d (:122-128).
 *   xm  : (batch, in_h, in_w, cin)  pre-modulated input  x[b,:,:,ci] * s[b,ci]   (or the plain input for EqualConv2d)
 *   wt  : (n_weight_taps, cout, cin)  shared weight, one K-major matrix per filter tap
 *   out : (batch, out_h, out_w, cout)
 * The convolution is described as up to 4 "phases" (1 for an ordinary convolution, 4 for the polyphase form of the
 * stride-2 transposed convolution).  Phase-local output position (m, n), m < rows, n < cols, reads input pixel
 * (m*in_stride + dy[t], n*in_stride + dx[t]) for each tap t (out-of-range pixels count as zero) with weight matrix
 * wt[widx[t]], and is written to output pixel (m*out_stride + out_y0, n*out_stride + out_x0).
 * Requirements: cin % 32 == 0, cout % 32 == 0 (tiles of 128 output channels; a partial tile rides on TMA zero fill),
 * pointers 16-byte aligned. */
typedef struct rick_conv_phase {
    int n_taps;
    int dy[9], dx[9], widx[9];
    int out_y0, out_x0;
    int rows, cols;
} rick_conv_phase;

typedef struct rick_conv_geom {
    int batch, in_h, in_w, cin, cout, out_h, out_w;
    int in_stride, out_stride;
    int n_weight_taps;
    int n_phases;
    rick_conv_phase phase[4];
} rick_conv_geom;

/* Fused epilogue (all pointers optional / NULL = skip):
 *   v = acc * demod[b,co] + noise_weight[0] * noise[b,oy,ox] + bias[co];  if (act) v = (v > 0 ? v : v*alpha) * scale
 *   out2 != NULL:  out = v;  out2 = v * s_next[b,co]   (the next layer's pre-modulated input, written in the same pass)
 *   out2 == NULL:  out = v * s_next[b,co]              (s_next == NULL counts as 1: plain output)
 * i.e. demodulate (model_probe_tune.py:249-251) + NoiseInjection (:293-298) + FusedLeakyReLU (op/fused_act.py). */
typedef struct rick_conv_epilogue {
    const float* demod;        /* (batch, cout) */
    const float* noise;        /* (batch, out_h, out_w) */
    const float* noise_weight; /* (1) on the device */
    const float* bias;         /* (cout) */
    const float* s_next;       /* (batch, cout) */
    void* out2;                /* (batch, out_h, out_w, cout) */
    int act;                   /* 0 = linear, 1 = leaky-ReLU */
    float alpha, scale;
} rick_conv_epilogue;

RICK_API int rick_conv_tc(void* out, const void* xm, const void* wt, const rick_conv_geom* geom,
                          const rick_conv_epilogue* epilogue, rick_stream_t stream);

/* The same kernel with the weight operand described by strides (in elements), so that ONE weight tensor in memory
 * serves both directions -- there is no transposed or re-packed copy:
 *   forward        GEMM-M = the layer's output channels, GEMM-K = its input channels.  For a (Cout, Cin, k, k) weight
 *                  stored channels-last ([Cout][k][k][Cin]): stride_m = k*k*Cin, stride_k = 1, stride_tap = Cin.
 *   data gradient  (autograd's cuDNN dgrad in the reference): geom describes the transposed problem -- geom->cout is the
 *                  layer's Cin, geom->cin its Cout, the taps are mirrored -- over the SAME memory: stride_m = 1,
 *                  stride_k = k*k*Cin, stride_tap = Cin.  stride_m == 1 makes the weight an MN-major UMMA operand.
 * Element (m, k, tap) of the operand is ptr[m*stride_m + k*stride_k + widx*stride_tap].  Exactly one of stride_m /
 * stride_k must be 1, the others multiples of 4 elements.  geom->cout % 32 == 0, geom->cin % 32 == 0.
 *
 * Split-K.  When the convolution has fewer output tiles than the GPU has SMs (the 4x4 ... 32x32 feature maps of the
 * batch-2 adaptation loop: a 512 -> 512 layer on a 4x4 map is 4 tiles, each streaming 2.4 MB of weights), the
 * (tap, input-channel-block) iterations of every tile are divided over several CTAs; their partial sums go to
 * `workspace` and a second kernel adds them in a fixed order (deterministic) and applies the epilogue.
 * rick_conv_tc_workspace(geom) returns the bytes that takes (0: no split for this geometry; negative: unsupported).
 * With workspace == NULL or workspace_bytes too small the call runs unsplit. */
typedef struct rick_conv_weight {
    const void* ptr;
    int64_t stride_m, stride_k, stride_tap;
} rick_conv_weight;

RICK_API int64_t rick_conv_tc_workspace(const rick_conv_geom* geom);
/* The launch plan rick_conv_tc_w would use for `geom` (host only, no device work): per phase the pixel tile
 * tile_w[i] x tile_h[i] (arrays of 4), the samples per tile, the K ranges per tile (1 = no split) and the number of output
 * tiles.  Tiles hold at most 256 pixels; single-phase launches are planned by waves over the SMs. */
RICK_API int rick_conv_tc_plan(const rick_conv_geom* geom, int* tile_w, int* tile_h, int* samples_per_tile, int* ksplit,
                               int* total_tiles);
RICK_API int rick_conv_tc_w(void* out, const void* xm, const rick_conv_weight* weight, const rick_conv_geom* geom,
                            const rick_conv_epilogue* epilogue, void* workspace, int64_t workspace_bytes,
                            rick_stream_t stream);

/* ------------------------------------------------------------------------------------------- tcgen05 weight gradient
 * dW of the convolutions above (in the reference: autograd's cuDNN wgrad behind model_probe_tune.py:122-128, 265, 274,
 * 280).  For each tap t < n_taps:
 *     dW[t][co][ci] = scale * sum_{b, m < rows, n < cols}  g[b, m*g_stride + gy[t], n*g_stride + gx[t], co]
 *                                                         * x[b, m*x_stride + xy[t], n*x_stride + xx[t], ci]
 * with g (batch, g_h, g_w, cout) and x (batch, x_h, x_w, cin) NHWC fp32; pixels outside either tensor count as zero.
 *   stride-1 / stride-s convolution (pad p):  (m, n) runs over the OUTPUT pixels; g_stride = 1, gy = gx = 0,
 *                                             x_stride = s, xy[t] = ky - p, xx[t] = kx - p
 *   stride-2 transposed convolution:          (m, n) runs over the INPUT pixels; x_stride = 1, xy = xx = 0,
 *                                             g_stride = 2, gy[t] = ky, gx[t] = kx
 * The result is written as dw[co*stride_co + ci*stride_ci + t*stride_tap] (element strides), i.e. directly in the
 * parameter's memory layout.  Partial sums over pixel ranges are folded in a fixed order: deterministic.
 * workspace: rick_conv_wgrad_workspace(geom) bytes (negative = unsupported geometry).  cout % 32 == 0, cin % 32 == 0. */
typedef struct rick_wgrad_geom {
    int batch;
    int g_h, g_w, cout;
    int x_h, x_w, cin;
    int n_taps;
    int gy[9], gx[9], xy[9], xx[9];
    int g_stride, x_stride;
    int rows, cols;
} rick_wgrad_geom;

RICK_API int64_t rick_conv_wgrad_workspace(const rick_wgrad_geom* geom);
RICK_API int rick_conv_wgrad_tc(void* dw, int64_t stride_co, int64_t stride_ci, int64_t stride_tap, const void* g,
                                const void* x, const rick_wgrad_geom* geom, void* workspace, float scale,
                                rick_stream_t stream);

/* ------------------------------------------------------------------------------------------- NHWC companions
 * rick_blur_nhwc: 4x4 FIR, up = down = 1, pads (pad0, pad1) on both axes, over x (batch, in_h, in_w, channels) fp32
 * -> out (batch, in_h+pad0+pad1-3, in_w+pad0+pad1-3, channels), with the rick_conv_epilogue fused (NULL = plain blur).
 * Replaces Blur after the transposed conv (model_probe_tune.py:268) + NoiseInjection + FusedLeakyReLU (StyledConv,
 * :342-348) in one pass.  channels % 4 == 0. */
RICK_API int rick_blur_nhwc(void* out, const void* x, const float* taps, int batch, int in_h, int in_w, int channels,
                            int pad0, int pad1, int flip_taps, const rick_conv_epilogue* epilogue,
                            rick_stream_t stream);

/* rick_to_rgb_nhwc: ToRGB (model_probe_tune.py:361-370).  rgb[b,o,y,x] (NCHW, 3 channels) =
 * sum_c y[b,y,x,c] * wmod[b,o,c] + bias[o] (+ skip[b,o,y,x] when skip != NULL), wmod = scale * W_rgb * style. */
RICK_API int rick_to_rgb_nhwc(float* rgb, const float* y, const float* wmod, const float* bias, const float* skip,
                              int batch, int h, int w, int channels, rick_stream_t stream);

/* ------------------------------------------------------------------------------------------- training-time fusions
 * Channels-last fp32 tensors (batch, hw, channels), channels % 4 == 0.  They replace the broadcast multiplies / adds and
 * ATen reductions autograd builds around the differentiated ModulatedConv2d (model_probe_tune.py:246-251), NoiseInjection
 * (:293-298) and FusedLeakyReLU (op/fused_act.py):
 *   rick_modulate_nhwc             y = x * s[b,c]
 *   rick_modulate_bwd_nhwc         gx = gy * s[b,c];  gs[b,c] = sum_hw gy * x
 *   rick_styled_epilogue_nhwc      y = lrelu(a * demod[b,c] + noise_weight * noise[b,hw] + bias[c]) * scale
 *   rick_styled_epilogue_bwd_nhwc  t = (y > 0 ? gy : gy*alpha) * scale;  ga = t * demod;  gdemod[b,c] = sum_hw t*a;
 *                                  gbias[c] = sum_{b,hw} t;  gnoise_weight = sum t * noise
 * workspace: rick_colsum_workspace(batch, hw, channels) bytes.  Reductions are deterministic (fixed-order folds). */
RICK_API int64_t rick_colsum_workspace(int batch, int64_t hw, int channels);
RICK_API int rick_modulate_nhwc(void* y, const void* x, const float* s, int batch, int64_t hw, int channels,
                                rick_stream_t stream);
RICK_API int rick_modulate_bwd_nhwc(void* gx, float* gs, void* workspace, const void* gy, const void* x, const float* s,
                                    int batch, int64_t hw, int channels, rick_stream_t stream);
RICK_API int rick_styled_epilogue_nhwc(void* y, const void* a, const float* demod, const float* noise,
                                       const float* noise_weight, const float* bias, int batch, int64_t hw,
                                       int channels, float alpha, float scale, rick_stream_t stream);
RICK_API int rick_styled_epilogue_bwd_nhwc(void* ga, float* gdemod, float* gbias, float* gnoise_weight, void* workspace,
                                           const void* gy, const void* y, const void* a, const float* demod,
                                           const float* noise, int batch, int64_t hw, int channels, float alpha,
                                           float scale, rick_stream_t stream);


/* ------------------------------------------------------------------------------------------- glue kernels
 * wsq[co][ci] = sum_tap w[co*stride_co + ci*stride_ci + tap*stride_tap]^2: the (Cout, Cin) table the demodulation of
 * ModulatedConv2d is computed from (model_probe_tune.py:249-251; ``w.pow(2).sum([2, 3])`` in one pass, no temporary). */
RICK_API int rick_weight_sqsum(float* out, const float* w, int cout, int cin, int taps, int64_t stride_co,
                               int64_t stride_ci, int64_t stride_tap, rick_stream_t stream);

/* Demodulation table of a ModulatedConv2d (model_probe_tune.py:246-251, algebraic form), batch <= 8:
 *     demod[b][co] = rsqrt(scale2 * sum_ci s[b][ci]^2 * wsq[co][ci] + eps);   s_out[b][ci] = s[b][ci] * s_scale (s_out may
 * be NULL).  rick_demod_bwd: first derivatives -- g_s[b][ci] = 2 s sum_co q wsq + s_scale g_sout (g_sout may be NULL),
 * g_wsq[co][ci] = sum_b q s^2 with q = -0.5 demod^3 scale2 g_demod; either output may be NULL. */
RICK_API int rick_demod_fwd(float* demod, float* s_out, const float* s, const float* wsq, int batch, int cin, int cout,
                            float scale2, float eps, float s_scale, rick_stream_t stream);
RICK_API int rick_demod_bwd(float* g_s, float* g_wsq, const float* g_demod, const float* g_sout, const float* demod,
                            const float* s, const float* wsq, int batch, int cin, int cout, float scale2, float s_scale,
                            rick_stream_t stream);

/* out = (a + b) * scale over n floats (n % 4 == 0, 16-byte aligned): the residual merge of ResBlock
 * (model_probe_tune.py:655-660, ``(out + skip) / sqrt(2)``) as one pass. */
RICK_API int rick_add_scale(float* out, const float* a, const float* b, float scale, int64_t n, rick_stream_t stream);

/* `count` small-batch EqualLinear layers in one launch (model_probe_tune.py:139-173: the 8 mapping-network layers, the
 * style -> channel modulation layer of every ModulatedConv2d):
 *     y[l][b, r] = act( w_scale[l] * sum_k w[l][r, k] * x[l][b, k] + b_scale[l] * bias[l][r] ),   b < batch <= 8
 * w[l] (out_dim[l], in_dim) row-major, x[l] rows x_stride[l] elements apart, y[l] (batch, out_dim[l]) contiguous,
 * bias[l] may be NULL.  act != 0: leaky-ReLU(alpha) * act_scale (fused_leaky_relu).  pixelnorm != 0: x is divided by
 * sqrt(mean_k x^2 + 1e-8) first (PixelNorm, :21-26).  Tables are HOST arrays.  in_dim % 4 == 0.
 * rick_linear_multi_wgrad: gw[l][r, k] = w_scale[l] * sum_b gy[l][b, r] * x[l][b, k] and
 * gbias[l][r] = b_scale[l] * sum_b gy[l][b, r] (either may be NULL per layer). */
RICK_API int rick_linear_multi(float* const* y, const float* const* w, const float* const* bias, const float* const* x,
                               const int64_t* x_stride, const int* out_dim, const float* w_scale, const float* b_scale,
                               int count, int batch, int in_dim, int act, float alpha, float act_scale, int pixelnorm,
                               rick_stream_t stream);
/* The discriminator's from-RGB layer, ConvLayer(3, C, 1) = 1x1 EqualConv2d + FusedLeakyReLU (model_probe_tune.py:595-641,
 * 676), as one pass each way.  img / gimg are NCHW (batch, cin <= 4, pixels), y / g NHWC (batch, pixels, cout), w (cout, cin).
 *   fwd       y = act( w_scale * sum_c img[c] * w[co,c] + bias[co] )        act != 0: leaky-ReLU(alpha) * act_scale
 *   bwd_data  gimg[c] = w_scale * sum_co w[co,c] * t[co],   t = act != 0 ? (y > 0 ? g : alpha*g) * act_scale : g */
RICK_API int rick_from_rgb_fwd(float* y, const float* img, const float* w, const float* bias, int batch, int64_t pixels,
                               int cin, int cout, float w_scale, int act, float alpha, float act_scale, rick_stream_t stream);
RICK_API int rick_from_rgb_bwd_data(float* gimg, const float* g, const float* y, const float* w, int batch, int64_t pixels,
                                    int cin, int cout, float w_scale, int act, float alpha, float act_scale,
                                    rick_stream_t stream);
RICK_API int rick_linear_multi_wgrad(float* const* gw, float* const* gbias, const float* const* gy, const float* const* x,
                                     const int64_t* x_stride, const int* out_dim, const float* w_scale,
                                     const float* b_scale, int count, int batch, int in_dim, rick_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RICK_B200_H */
