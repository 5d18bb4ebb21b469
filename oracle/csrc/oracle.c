/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the order-sensitive pieces of the RICK hot path.
 *
 * Nothing in the product (rick_b200/) links or loads this file.  It is built by
 * oracle/Makefile into oracle/_build/liboracle.so and used by tests/ as an
 * independent checker next to the torch/NumPy oracle:
 *
 *   oracle_upfirdn2d        direct-form statement of the operator the reference computes in
 *                           op/upfirdn2d.py:159-200 (zero-stuff, pad/crop, correlate with flipped
 *                           taps, decimate), one output sample at a time, double accumulation.
 *   oracle_bias_act         op/fused_bias_act_kernel.cu:26-47 arithmetic (act 1/3, grad 0/1/2).
 *   oracle_row_mean_f32     NumPy's float32 ``ndarray.mean`` over contiguous rows: pairwise
 *                           summation (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum,
 *                           NumPy 2.3.x: blocks of 128, 8 partial sums) followed by a float32
 *                           divide by the row length -- train_dynamic_update_prune.py:282, 291, 339.
 *   oracle_percentile_linear np.percentile(v, q) with the default 'linear' method on a float64
 *                           vector (numpy/lib/_function_base_impl.py, _quantile/_lerp) --
 *                           train_dynamic_update_prune.py:285-286, 298-299, 352-353.
 *   oracle_decide           freeze / fine-tune / prune decisions, train:312-314 (strict) and
 *                           382-384 (D skip variant).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline long floordiv(long a, long b) { long q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

/* x: (planes, in_h, in_w)  taps: (kh, kw)  out: (planes, out_h, out_w), all float32 */
int oracle_upfirdn2d(const float* x, const float* taps, float* out, long planes, long in_h, long in_w,
                     int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                     int pad_x0, int pad_x1, int pad_y0, int pad_y1) {
    long out_h = floordiv(in_h * up_y + pad_y0 + pad_y1 - kh, down_y) + 1;
    long out_w = floordiv(in_w * up_x + pad_x0 + pad_x1 - kw, down_x) + 1;
    if (out_h <= 0 || out_w <= 0) return 1;
    for (long p = 0; p < planes; ++p)
        for (long oy = 0; oy < out_h; ++oy)
            for (long ox = 0; ox < out_w; ++ox) {
                double acc = 0.0;
                for (int ty = 0; ty < kh; ++ty) {
                    /* position in the zero-stuffed, padded signal touched by flipped tap ty */
                    long uy = oy * down_y + ty - pad_y0;
                    if (uy < 0 || uy % up_y) continue;
                    long iy = uy / up_y;
                    if (iy >= in_h) continue;
                    for (int tx = 0; tx < kw; ++tx) {
                        long ux = ox * down_x + tx - pad_x0;
                        if (ux < 0 || ux % up_x) continue;
                        long ix = ux / up_x;
                        if (ix >= in_w) continue;
                        acc += (double)x[(p * in_h + iy) * in_w + ix] * (double)taps[(kh - 1 - ty) * kw + (kw - 1 - tx)];
                    }
                }
                out[(p * out_h + oy) * out_w + ox] = (float)acc;
            }
    return 0;
}

/* x/ref/out flat (n), bias (size_b) indexed by (i / step_b) % size_b; bias/ref may be NULL */
int oracle_bias_act(const float* x, const float* bias, const float* ref, float* out, long n, long step_b,
                    long size_b, int act, int grad, float alpha, float scale) {
    for (long i = 0; i < n; ++i) {
        float v = x[i];
        if (bias) v += bias[(i / step_b) % size_b];
        float r = ref ? ref[i] : 0.0f;
        float y;
        if (grad == 2) y = 0.0f;
        else if (act == 3) y = ((grad == 0 ? v : r) > 0.0f) ? v : v * alpha;
        else y = v;
        out[i] = y * scale;
    }
    return 0;
}

/* NumPy float32 pairwise sum over n contiguous elements */
static float pairwise_sum_f32(const float* a, long n) {
    if (n < 8) {
        float res = 0.0f;            /* numpy starts from 0. (the -0.0 nuance is irrelevant for sums of squares) */
        for (long i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        long i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    long n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sum_f32(a, n2) + pairwise_sum_f32(a + n2, n - n2);
}

/* fim[r] = mean(row r)            (bias == NULL)
 * fim[r] = (mean(row r) + bias[r]) / 2   otherwise;   all float32 arithmetic */
int oracle_row_mean_f32(const float* a, const float* bias, float* fim, long rows, long len) {
    for (long r = 0; r < rows; ++r) {
        float m = pairwise_sum_f32(a + r * len, len) / (float)len;
        fim[r] = bias ? (m + bias[r]) / 2.0f : m;
    }
    return 0;
}

static int cmp_f64(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* np.percentile(v, q), method='linear', v float64 of length n (copied and sorted here) */
double oracle_percentile_linear(const double* v, long n, double q) {
    double* s = (double*)malloc(sizeof(double) * (size_t)n);
    memcpy(s, v, sizeof(double) * (size_t)n);
    qsort(s, (size_t)n, sizeof(double), cmp_f64);
    double quant = q / 100.0;
    double virt = (double)(n - 1) * quant;
    double prev = floor(virt);
    long lo = (long)prev;
    long hi = lo + 1;
    if (lo < 0) lo = 0;
    if (lo > n - 1) lo = n - 1;
    if (hi > n - 1) hi = n - 1;
    double gamma = virt - prev;
    double a = s[lo], b = s[hi];
    double diff = b - a;
    double res = (gamma >= 0.5) ? b - diff * (1.0 - gamma) : a + diff * gamma;
    free(s);
    return res;
}

/* state[i] bits: 1 = freeze (fim > cut), 2 = prune, 4 = fine-tune; each evaluated independently, exactly as the
 * three np.where() calls of train:312-314.  closed_low selects the D-skip comparisons (train:382-384). */
int oracle_decide(const float* fim, long n, double cut, double prune, int closed_low, uint8_t* state) {
    for (long i = 0; i < n; ++i) {
        double f = (double)fim[i];
        uint8_t s = 0;
        if (f > cut) s |= 1;
        if (closed_low ? (f < prune) : (f <= prune)) s |= 2;
        if ((closed_low ? (f >= prune) : (f > prune)) && f <= cut) s |= 4;
        state[i] = s;
    }
    return 0;
}
