/*
 * f8_oracle.c -- CPU restatement of F8Net's int_op_only forward primitives.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library, and only as the checker or the
 * CPU baseline.  The product path (f8net_b200/) never links or calls it.
 *
 * Parity pinning: the reference ships no golden vectors or tests
 * (SURVEY.md section 4), and the integer tensor arithmetic it calls lives in
 * third-party PyTorch (pinned torch==1.11.0, /root/reference/requirements.txt:28;
 * not vendored).  This restatement is therefore pinned by executing the
 * reference's own source (oracle/ref_harness.py, in the authoring container,
 * on torch 2.11) and comparing every layer's accumulator and the final logits
 * on seeded fixtures; the resulting vectors are committed under tests/golden/
 * together with tests/golden/make_golden.py.
 *
 * All tensors are int32, NCHW, contiguous -- the reference's layout.
 * Every add / multiply / left shift wraps mod 2^32 (torch int32 semantics),
 * implemented with uint32_t arithmetic so the C is free of signed overflow UB.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define F8O_API __attribute__((visibility("default")))

static inline int32_t wrap_add(int32_t a, int32_t b) {
    return (int32_t)((uint32_t)a + (uint32_t)b);
}
static inline int32_t wrap_shl(int32_t a, int s) {
    return (int32_t)((uint32_t)a << s);
}
/* arithmetic right shift (torch __rshift__ on int32) */
static inline int32_t asr(int32_t a, int s) {
    return a >= 0 ? (a >> s) : ~((~a) >> s);
}

F8O_API void f8o_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

F8O_API int f8o_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * int_op_only_fix_quant(input, wl=8, fl, input_fl, signed)
 * /root/reference/models/fix_quant_ops.py:90-114
 *   net_fl = input_fl - fl
 *   net_fl > 0 : res = input + (1 << (net_fl-1))                       (:102, wraps)
 *                tie  = (input % (1<<net_fl)) == (1 << (net_fl-1))     (:103, floor-mod)
 *                res  = tie ? ((res >> (net_fl+1)) << 1) : res >> net_fl   (:103-104)
 *   else       : res = input << (-net_fl)                              (:106, wraps)
 *   signed     : clamp [-127, 127]  else clamp [0, 255]                (:107-112)
 * Scalar version; element-wise over n values.
 */
static inline int32_t requant1(int32_t x, int net_fl, int is_signed) {
    int32_t res;
    if (net_fl > 0) {
        const int32_t half = (int32_t)1 << (net_fl - 1);
        const uint32_t mask = ((uint32_t)1 << net_fl) - 1u;
        int32_t r = wrap_add(x, half);
        /* python/torch remainder with positive divisor == low bits of two's complement */
        if (((uint32_t)x & mask) == (uint32_t)half)
            res = wrap_shl(asr(r, net_fl + 1), 1);
        else
            res = asr(r, net_fl);
    } else {
        res = wrap_shl(x, -net_fl);
    }
    if (is_signed) {
        if (res > 127) res = 127;
        if (res < -127) res = -127;
    } else {
        if (res > 255) res = 255;
        if (res < 0) res = 0;
    }
    return res;
}

F8O_API int f8o_requant(const int32_t *x, int32_t *y, size_t n, int fl,
                        int input_fl, int is_signed) {
    int net_fl = input_fl - fl;
    if (net_fl > 30 || net_fl < -31) return -1; /* outside what torch can express */
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i)
        y[i] = requant1(x[i], net_fl, is_signed);
    return 0;
}

/*
 * Integer nn.Conv2d built by ReLUClipFXQConvBN.int_conv()
 * /root/reference/models/fix_quant_ops.py:680-714, called at
 * fix_resnet.py:34,59,356 ; fix_mobilenet_v1.py:33,123 ; fix_mobilenet_v2.py:28,210,221.
 * torch (third party, not vendored) evaluates it as
 *   y[n,o,p,q] = b[o] + sum_{c,r,s} x[n, g*Cg + c, p*st + r - pad, q*st + s - pad] * w[o,c,r,s]
 * with zero padding and int32 wrap-around.  Restated here as a direct loop nest.
 *   x [N,C,H,W]  w [O, C/groups, kh, kw]  b [O]  y [N,O,Ho,Wo]
 */
F8O_API int f8o_conv2d(const int32_t *x, int N, int C, int H, int W,
                       const int32_t *w, const int32_t *b, int O, int kh,
                       int kw, int stride, int pad, int groups, int32_t *y) {
    if (C % groups || O % groups) return -1;
    const int Cg = C / groups, Og = O / groups;
    const int Ho = (H + 2 * pad - kh) / stride + 1;
    const int Wo = (W + 2 * pad - kw) / stride + 1;
    if (stride == 1) {
        /* Same sum, evaluated on a zero-padded copy so that every tap is one long
         * contiguous multiply-add over the (padded-width) output plane: keeps the
         * checker fast on the 7x7 / 14x14 stages.  Integer adds commute mod 2^32, so
         * the summation order does not matter. */
        const int Hp = H + 2 * pad, Wp = W + 2 * pad;
        const size_t plane = (size_t)Hp * Wp;
        int32_t *xp = (int32_t *)x;
        if (pad) {
            xp = (int32_t *)calloc((size_t)N * C * plane + (size_t)kw, sizeof(int32_t));
            if (!xp) return -3;
#pragma omp parallel for schedule(static)
            for (ptrdiff_t nc = 0; nc < (ptrdiff_t)N * C; ++nc)
                for (int h = 0; h < H; ++h)
                    memcpy(xp + (size_t)nc * plane + (size_t)(h + pad) * Wp + pad,
                           x + ((size_t)nc * H + h) * W, (size_t)W * sizeof(int32_t));
        }
        const size_t run = (size_t)(Ho - 1) * Wp + Wo; /* valid span in padded-width coords */
#pragma omp parallel
        {
            uint32_t *acc = (uint32_t *)malloc(((size_t)Ho * Wp) * sizeof(uint32_t));
#pragma omp for collapse(2) schedule(dynamic)
            for (int n = 0; n < N; ++n) {
                for (int o = 0; o < O; ++o) {
                    const int g = o / Og;
                    const uint32_t bias = b ? (uint32_t)b[o] : 0u;
                    for (size_t i = 0; i < run; ++i) acc[i] = bias;
                    for (int c = 0; c < Cg; ++c) {
                        const uint32_t *xc = (const uint32_t *)xp +
                                             ((size_t)n * C + (size_t)g * Cg + c) * plane;
                        for (int r = 0; r < kh; ++r)
                            for (int s = 0; s < kw; ++s) {
                                const uint32_t wv =
                                    (uint32_t)w[(((size_t)o * Cg + c) * kh + r) * kw + s];
                                if (wv == 0u) continue;
                                const uint32_t *xs = xc + (size_t)r * Wp + s;
                                for (size_t i = 0; i < run; ++i) acc[i] += xs[i] * wv;
                            }
                    }
                    int32_t *yo = y + ((size_t)n * O + o) * Ho * Wo;
                    for (int p = 0; p < Ho; ++p)
                        memcpy(yo + (size_t)p * Wo, acc + (size_t)p * Wp,
                               (size_t)Wo * sizeof(int32_t));
                }
            }
            free(acc);
        }
        if (pad) free(xp);
        return 0;
    }
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int n = 0; n < N; ++n) {
        for (int o = 0; o < O; ++o) {
            const int g = o / Og;
            uint32_t *yo = (uint32_t *)(y + ((size_t)n * O + o) * Ho * Wo);
            const uint32_t bias = b ? (uint32_t)b[o] : 0u;
            for (int i = 0; i < Ho * Wo; ++i) yo[i] = bias;
            for (int c = 0; c < Cg; ++c) {
                const int32_t *xc = x + ((size_t)n * C + (size_t)g * Cg + c) * H * W;
                for (int r = 0; r < kh; ++r) {
                    for (int s = 0; s < kw; ++s) {
                        const uint32_t wv =
                            (uint32_t)w[(((size_t)o * Cg + c) * kh + r) * kw + s];
                        if (wv == 0u) continue;
                        /* valid output range for this tap */
                        int p_lo = 0, p_hi = Ho, q_lo = 0, q_hi = Wo;
                        while (p_lo < Ho && p_lo * stride + r - pad < 0) ++p_lo;
                        while (p_hi > p_lo && (p_hi - 1) * stride + r - pad >= H) --p_hi;
                        while (q_lo < Wo && q_lo * stride + s - pad < 0) ++q_lo;
                        while (q_hi > q_lo && (q_hi - 1) * stride + s - pad >= W) --q_hi;
                        for (int p = p_lo; p < p_hi; ++p) {
                            const int32_t *xr = xc + (size_t)(p * stride + r - pad) * W + (s - pad);
                            uint32_t *yr = yo + (size_t)p * Wo;
                            for (int q = q_lo; q < q_hi; ++q)
                                yr[q] += (uint32_t)xr[q * stride] * wv;
                        }
                    }
                }
            }
        }
    }
    return 0;
}

/* nn.ReLU on int32 (IntBlock.body / post_relu / head[1] / tail[1]); in place allowed */
F8O_API void f8o_relu(const int32_t *x, int32_t *y, size_t n) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) y[i] = x[i] > 0 ? x[i] : 0;
}

/* float32 -> int32 the way the reference's x86 CPU path does it (.int() on a float
 * tensor): truncation; out-of-range yields the x86 "integer indefinite" 0x80000000. */
static inline int32_t f2i_x86(float f) {
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return INT32_MIN;
    return (int32_t)f;
}

/*
 * ResNet head max-pool with quant_maxpool False:
 *   x = self.head[-1](x.float()).int()      /root/reference/models/fix_resnet.py:358-359
 * head[-1] = nn.MaxPool2d(3, 2, 1)          fix_resnet.py:439
 * int32 -> float32 (round to nearest even) -> window max with -inf padding -> trunc to int32.
 */
F8O_API int f8o_maxpool_float_rt(const int32_t *x, int N, int C, int H, int W,
                                 int k, int stride, int pad, int32_t *y) {
    const int Ho = (H + 2 * pad - k) / stride + 1;
    const int Wo = (W + 2 * pad - k) / stride + 1;
#pragma omp parallel for schedule(static)
    for (ptrdiff_t nc = 0; nc < (ptrdiff_t)N * C; ++nc) {
        const int32_t *xc = x + (size_t)nc * H * W;
        int32_t *yc = y + (size_t)nc * Ho * Wo;
        for (int p = 0; p < Ho; ++p)
            for (int q = 0; q < Wo; ++q) {
                float m = -INFINITY;
                for (int r = 0; r < k; ++r) {
                    int h = p * stride + r - pad;
                    if (h < 0 || h >= H) continue;
                    for (int s = 0; s < k; ++s) {
                        int ww = q * stride + s - pad;
                        if (ww < 0 || ww >= W) continue;
                        float f = (float)xc[(size_t)h * W + ww];
                        if (f > m) m = f;
                    }
                }
                yc[(size_t)p * Wo + q] = f2i_x86(m);
            }
    }
    return 0;
}

/*
 * FXQMaxPool2d.forward  /root/reference/models/fix_quant_ops.py:141-157
 * (only with quant_maxpool True, which no shipped int_op_only config sets):
 * zero padding + unfold + integer max.
 */
F8O_API int f8o_maxpool_int(const int32_t *x, int N, int C, int H, int W, int k,
                            int stride, int pad, int32_t *y) {
    const int Ho = (H + 2 * pad - k) / stride + 1;
    const int Wo = (W + 2 * pad - k) / stride + 1;
#pragma omp parallel for schedule(static)
    for (ptrdiff_t nc = 0; nc < (ptrdiff_t)N * C; ++nc) {
        const int32_t *xc = x + (size_t)nc * H * W;
        int32_t *yc = y + (size_t)nc * Ho * Wo;
        for (int p = 0; p < Ho; ++p)
            for (int q = 0; q < Wo; ++q) {
                int32_t m = INT32_MIN;
                for (int r = 0; r < k; ++r)
                    for (int s = 0; s < k; ++s) {
                        int h = p * stride + r - pad, ww = q * stride + s - pad;
                        int32_t v = (h < 0 || h >= H || ww < 0 || ww >= W)
                                        ? 0 : xc[(size_t)h * W + ww];
                        if (v > m) m = v;
                    }
                yc[(size_t)p * Wo + q] = m;
            }
    }
    return 0;
}

/*
 * FXQAvgPool2d.forward, int_op_only branch
 * /root/reference/models/fix_quant_ops.py:126-134
 *   res = x.sum(-1).sum(-1)   (int64)  ; assert res <= 2^32-1 ; res = res.int() (wraps)
 * Returns -2 when the reference's assert would fire.
 *   x [N,C,H,W] -> y [N,C]
 */
F8O_API int f8o_avgpool_sum(const int32_t *x, int N, int C, int H, int W,
                            int32_t *y) {
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (ptrdiff_t nc = 0; nc < (ptrdiff_t)N * C; ++nc) {
        const int32_t *xc = x + (size_t)nc * H * W;
        int64_t acc = 0;
        for (int i = 0; i < H * W; ++i) acc += xc[i];
        if (acc > (int64_t)4294967295LL) bad |= 1;
        y[nc] = (int32_t)(uint32_t)(uint64_t)acc;
    }
    return bad ? -2 : 0;
}

/*
 * Integer nn.Linear built by ReLUClipFXQLinear.int_fc()
 * /root/reference/models/fix_quant_ops.py:1165-1195, called at fix_resnet.py:383,
 * fix_mobilenet_v1.py:147, fix_mobilenet_v2.py:241, followed by .float().
 *   q [N,K]  w [O,K]  b [O]  ->  y_int [N,O] (may be NULL), y_float [N,O] (may be NULL)
 */
F8O_API int f8o_linear(const int32_t *q, int N, int K, const int32_t *w,
                       const int32_t *b, int O, int32_t *y_int, float *y_float) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
        for (int o = 0; o < O; ++o) {
            uint32_t acc = b ? (uint32_t)b[o] : 0u;
            const int32_t *qr = q + (size_t)n * K, *wr = w + (size_t)o * K;
            for (int k = 0; k < K; ++k) acc += (uint32_t)qr[k] * (uint32_t)wr[k];
            if (y_int) y_int[(size_t)n * O + o] = (int32_t)acc;
            if (y_float) y_float[(size_t)n * O + o] = (float)(int32_t)acc;
        }
    return 0;
}

/*
 * Residual align + add + clamp
 * /root/reference/models/fix_resnet.py:40-76 (strict '>'), fix_mobilenet_v2.py:34-48 ('>=';
 * no behavioural difference because a shift by 0 is the identity).
 *   if res_fl > x_fl : x <<= (res_fl - x_fl) else res <<= (x_fl - res_fl)      (wrap)
 *   res += x (wrap) ; clamp(-2^31+1, 2^31-1)
 * Writes into out (may alias res).  Returns the output fraclen = max(res_fl, x_fl).
 */
F8O_API int f8o_residual_add(const int32_t *res, const int32_t *x, int32_t *out,
                             size_t n, int res_fl, int x_fl) {
    const int d = res_fl - x_fl;
    if (d > 31 || d < -31) return -1;
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) {
        int32_t r = res[i], s = x[i];
        if (d > 0) s = wrap_shl(s, d); else r = wrap_shl(r, -d);
        int32_t v = wrap_add(r, s);
        if (v < -2147483647) v = -2147483647;
        out[i] = v;
    }
    return d > 0 ? res_fl : x_fl;
}

/*
 * forward_loss input integerisation, normalize False branch
 * /root/reference/fix_train.py:689-692 :  input = (255 * input).round_().int()
 * torch.round_ is round-half-to-even == nearbyintf in the default rounding mode.
 */
F8O_API void f8o_input_u8(const float *x, int32_t *y, size_t n) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i)
        y[i] = f2i_x86(nearbyintf(255.0f * x[i]));
}

/*
 * forward_loss input integerisation, normalize True branch
 * /root/reference/fix_train.py:682-687 with fix_quant (fix_quant_ops.py:64-87), signed:
 *   res = round(x * 2^fl) ; clamp +-127 ; res /= 2^fl ; input = (res * 2^fl).int()
 * All steps are exact in float32 (power-of-two scaling of an integer |v| <= 127).
 */
F8O_API void f8o_input_s8(const float *x, int32_t *y, size_t n, int fl) {
    const float sc = ldexpf(1.0f, fl);
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) {
        float r = nearbyintf(x[i] * sc);
        if (r > 127.0f) r = 127.0f;
        if (r < -127.0f) r = -127.0f;
        r = r / sc;
        y[i] = f2i_x86(r * sc);
    }
}
