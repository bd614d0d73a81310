/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the RubiksNet shift kernels, instantiated
 * once per scalar type by rubiks_oracle.c (REAL = float, double; SFX = _f32, _f64).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this.  The product path (rubiksnet_b200/) never does.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 * Layouts: 3D tensors are contiguous [N,T,C,H,W], shift is [3,C] rows (T,H,W)
 * (cuda_src/rubiks.cpp:243-244); 2D tensors are [N,C,H,W], shift [2,C] rows (H,W).
 *
 * Deliberate differences from the reference, none of which change a defined result:
 *   - the per-(axis,channel) shift-gradient sum is accumulated in double and rounded once, instead of
 *     REAL atomicAdd in arbitrary order (cuda_src/rubiks3d_kernels.cu:448-450 + rubiks.cpp:344-345);
 *   - the 2D kernels' uint32_t stride/pad arithmetic (cuda_src/rubiks2d_kernels.cu:98-107,298-305)
 *     is written with int: for stride>0 the wrap-around tests select exactly the same taps.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

/* floorf() on the shift even for the double instantiation: cuda_src/rubiks3d_kernels.cu:65-69 */
static inline int FN(floor3d)(REAL s) { return (int)floorf((float)s); }

/* zero-padded read of x[n,t,c,h,w]: the bounds test of cuda_src/rubiks3d_kernels.cu:102-106 */
static inline REAL FN(tap3d)(const REAL *x, int64_t plane_base, int t, int h, int w, int T, int C,
                            int H, int W) {
    if (t < 0 || h < 0 || w < 0 || t >= T || h >= H || w >= W) return (REAL)0;
    return x[plane_base + (int64_t)t * C * H * W + (int64_t)h * W + w];
}

/* cuda_src/rubiks3d_kernels.cu:15-205 (forward, incl. the quantize branch :76-93) */
void FN(oracle_shift3d_forward)(const REAL *x, const REAL *shift, REAL *out, int N, int T, int C,
                                int H, int W, int To, int Ho, int Wo, int sT, int sH, int sW, int pT,
                                int pH, int pW, int quantize) {
    const int64_t HW = (int64_t)H * W, HWo = (int64_t)Ho * Wo;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
        for (int c = 0; c < C; ++c) {
            const REAL st = shift[c], sh = shift[C + c], sw = shift[2 * C + c];
            const int ft = FN(floor3d)(st), fh = FN(floor3d)(sh), fw = FN(floor3d)(sw);
            const REAL rt = st - ft, rh = sh - fh, rw = sw - fw;
            /* base of x[n, 0, c, 0, 0]; tap3d adds t*C*HW */
            const int64_t xb = ((int64_t)n * T * C + c) * HW;
            for (int to = 0; to < To; ++to)
                for (int ho = 0; ho < Ho; ++ho)
                    for (int wo = 0; wo < Wo; ++wo) {
                        const int bt = to * sT - pT, bh = ho * sH - pH, bw = wo * sW - pW;
                        REAL v;
                        if (quantize) {
                            const int kt = (rt < 0.5f) ? ft : ft + 1;
                            const int kh = (rh < 0.5f) ? fh : fh + 1;
                            const int kw = (rw < 0.5f) ? fw : fw + 1;
                            v = FN(tap3d)(x, xb, bt + kt, bh + kh, bw + kw, T, C, H, W);
                        } else {
                            const int t0 = bt + ft, h0 = bh + fh, w0 = bw + fw;
                            const REAL q111 = FN(tap3d)(x, xb, t0, h0, w0, T, C, H, W);
                            const REAL q112 = FN(tap3d)(x, xb, t0, h0, w0 + 1, T, C, H, W);
                            const REAL q121 = FN(tap3d)(x, xb, t0, h0 + 1, w0, T, C, H, W);
                            const REAL q122 = FN(tap3d)(x, xb, t0, h0 + 1, w0 + 1, T, C, H, W);
                            const REAL q211 = FN(tap3d)(x, xb, t0 + 1, h0, w0, T, C, H, W);
                            const REAL q212 = FN(tap3d)(x, xb, t0 + 1, h0, w0 + 1, T, C, H, W);
                            const REAL q221 = FN(tap3d)(x, xb, t0 + 1, h0 + 1, w0, T, C, H, W);
                            const REAL q222 = FN(tap3d)(x, xb, t0 + 1, h0 + 1, w0 + 1, T, C, H, W);
                            /* association order of :193-203 */
                            v = (1 - rt) * ((1 - rh) * (q111 * (1 - rw) + q112 * rw) +
                                            rh * (q121 * (1 - rw) + q122 * rw)) +
                                rt * ((1 - rh) * (q211 * (1 - rw) + q212 * rw) +
                                      rh * (q221 * (1 - rw) + q222 * rw));
                        }
                        out[(((int64_t)n * To + to) * C + c) * HWo + (int64_t)ho * Wo + wo] = v;
                    }
        }
}

/* cuda_src/rubiks3d_kernels.cu:208-215 */
static inline REAL FN(interp2)(REAL p11, REAL p12, REAL p21, REAL p22, REAL d1, REAL d2) {
    return p11 * (1 - d1) * (1 - d2) + p12 * (1 - d1) * d2 + p21 * d1 * (1 - d2) + p22 * d1 * d2;
}

/*
 * cuda_src/rubiks3d_kernels.cu:218-452 (per-pixel shift gradient incl. the exact-integer "a" taps
 * :290-298,359-431) followed by the reduction of rubiks.cpp:295-299,344-345.  Writes the
 * UN-normalised gradient [3,C] (addmv_ with beta=0 overwrites).  `quantize` is ignored by the
 * reference here, so there is no such argument.
 */
void FN(oracle_shift3d_backward_shift)(const REAL *x, const REAL *shift, const REAL *og,
                                       REAL *gshift, int N, int T, int C, int H, int W, int To,
                                       int Ho, int Wo, int sT, int sH, int sW, int pT, int pH,
                                       int pW) {
    const int64_t HW = (int64_t)H * W, HWo = (int64_t)Ho * Wo;
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < C; ++c) {
        const REAL st = shift[c], sh = shift[C + c], sw = shift[2 * C + c];
        const int ft = FN(floor3d)(st), fh = FN(floor3d)(sh), fw = FN(floor3d)(sw);
        const REAL rt = st - ft, rh = sh - fh, rw = sw - fw;
        const int at = (rt == 0) ? -1 : 0, ah = (rh == 0) ? -1 : 0, aw = (rw == 0) ? -1 : 0;
        double accT = 0, accH = 0, accW = 0;
        for (int n = 0; n < N; ++n) {
            const int64_t xb = ((int64_t)n * T * C + c) * HW;
            for (int to = 0; to < To; ++to)
                for (int ho = 0; ho < Ho; ++ho)
                    for (int wo = 0; wo < Wo; ++wo) {
                        const int bt = to * sT - pT, bh = ho * sH - pH, bw = wo * sW - pW;
                        /* "small := small_a" on every axis whose remainder is exactly 0, "large" unchanged */
                        const int tl = bt + ft + at, th = bt + ft + 1;
                        const int hl = bh + fh + ah, hh = bh + fh + 1;
                        const int wl = bw + fw + aw, wh = bw + fw + 1;
                        const REAL q111 = FN(tap3d)(x, xb, tl, hl, wl, T, C, H, W);
                        const REAL q112 = FN(tap3d)(x, xb, tl, hl, wh, T, C, H, W);
                        const REAL q121 = FN(tap3d)(x, xb, tl, hh, wl, T, C, H, W);
                        const REAL q122 = FN(tap3d)(x, xb, tl, hh, wh, T, C, H, W);
                        const REAL q211 = FN(tap3d)(x, xb, th, hl, wl, T, C, H, W);
                        const REAL q212 = FN(tap3d)(x, xb, th, hl, wh, T, C, H, W);
                        const REAL q221 = FN(tap3d)(x, xb, th, hh, wl, T, C, H, W);
                        const REAL q222 = FN(tap3d)(x, xb, th, hh, wh, T, C, H, W);
                        const REAL gT = -FN(interp2)(q111, q112, q121, q122, rh, rw) +
                                        FN(interp2)(q211, q212, q221, q222, rh, rw);
                        const REAL gH = -FN(interp2)(q111, q112, q211, q212, rt, rw) +
                                        FN(interp2)(q121, q122, q221, q222, rt, rw);
                        const REAL gW = -FN(interp2)(q111, q121, q211, q221, rt, rh) +
                                        FN(interp2)(q112, q122, q212, q222, rt, rh);
                        const REAL up =
                            og[(((int64_t)n * To + to) * C + c) * HWo + (int64_t)ho * Wo + wo];
                        accT += (double)(REAL)(gT * up);
                        accH += (double)(REAL)(gH * up);
                        accW += (double)(REAL)(gW * up);
                    }
        }
        gshift[c] = (REAL)accT;
        gshift[C + c] = (REAL)accH;
        gshift[2 * C + c] = (REAL)accW;
    }
}

/* cuda_src/rubiks3d_kernels.cu:932-960: nothing is written when the norm is 0 */
void FN(oracle_normalize_shift_grad_3d)(REAL *g, int C, REAL factor) {
    for (int c = 0; c < C; ++c) {
        REAL gt, gh, gw;
        if (factor < 0) {
            gt = g[c];
            gh = 0;
            gw = 0;
        } else {
            gt = g[c] * factor;
            gh = g[C + c];
            gw = g[2 * C + c];
        }
        const REAL mag = (REAL)sqrt((double)(gt * gt + gh * gh + gw * gw));
        if (mag > 0) {
            g[c] = gt / mag;
            g[C + c] = gh / mag;
            g[2 * C + c] = gw / mag;
        }
    }
}

/* zero-padded, stride-aware read of og used by the adjoint: cuda_src/rubiks3d_kernels.cu:586-594 */
static inline REAL FN(tap3d_adj)(const REAL *og, int64_t plane_base, int t, int h, int w, int sT,
                                int sH, int sW, int To, int C, int Ho, int Wo) {
    if (t % sT != 0 || h % sH != 0 || w % sW != 0) return (REAL)0; /* C '%': truncating */
    t /= sT;
    h /= sH;
    w /= sW;
    if (t < 0 || h < 0 || w < 0 || t >= To || h >= Ho || w >= Wo) return (REAL)0;
    return og[plane_base + (int64_t)t * C * Ho * Wo + (int64_t)h * Wo + w];
}

/*
 * cuda_src/rubiks3d_kernels.cu:455-723 (general) and :726-929 (stride-1/pad-0 twin; same arithmetic
 * with stride 1, pad 0), dispatcher :1112-1144.
 */
void FN(oracle_shift3d_backward_input)(const REAL *shift, const REAL *og, REAL *gin, int N, int T,
                                       int C, int H, int W, int To, int Ho, int Wo, int sT, int sH,
                                       int sW, int pT, int pH, int pW, int quantize) {
    const int64_t HW = (int64_t)H * W, HWo = (int64_t)Ho * Wo;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
        for (int c = 0; c < C; ++c) {
            const REAL st = -shift[c], sh = -shift[C + c], sw = -shift[2 * C + c];
            const int ft = FN(floor3d)(st), fh = FN(floor3d)(sh), fw = FN(floor3d)(sw);
            const REAL rt = st - ft, rh = sh - fh, rw = sw - fw;
            const int64_t gb = ((int64_t)n * To * C + c) * HWo;
            for (int t = 0; t < T; ++t)
                for (int h = 0; h < H; ++h)
                    for (int w = 0; w < W; ++w) {
                        const int bt = t + pT, bh = h + pH, bw = w + pW;
                        REAL v;
#define ADJ(tt, hh, ww) FN(tap3d_adj)(og, gb, tt, hh, ww, sT, sH, sW, To, C, Ho, Wo)
                        if (quantize) { /* :533-558 */
                            const int kt = (rt < 0.5f) ? ft : ft + 1;
                            const int kh = (rh < 0.5f) ? fh : fh + 1;
                            const int kw = (rw < 0.5f) ? fw : fw + 1;
                            v = ADJ(bt + kt, bh + kh, bw + kw);
                        } else if (st == 0 && sh == 0 && sw == 0) { /* :561-576 */
                            v = ADJ(bt, bh, bw);
                        } else { /* :582-719 */
                            const int t0 = bt + ft, h0 = bh + fh, w0 = bw + fw;
                            const REAL q111 = ADJ(t0, h0, w0), q112 = ADJ(t0, h0, w0 + 1);
                            const REAL q121 = ADJ(t0, h0 + 1, w0), q122 = ADJ(t0, h0 + 1, w0 + 1);
                            const REAL q211 = ADJ(t0 + 1, h0, w0), q212 = ADJ(t0 + 1, h0, w0 + 1);
                            const REAL q221 = ADJ(t0 + 1, h0 + 1, w0),
                                       q222 = ADJ(t0 + 1, h0 + 1, w0 + 1);
                            v = (1 - rt) * ((1 - rh) * (q111 * (1 - rw) + q112 * rw) +
                                            rh * (q121 * (1 - rw) + q122 * rw)) +
                                rt * ((1 - rh) * (q211 * (1 - rw) + q212 * rw) +
                                      rh * (q221 * (1 - rw) + q222 * rw));
                        }
#undef ADJ
                        gin[(((int64_t)n * T + t) * C + c) * HW + (int64_t)h * W + w] = v;
                    }
        }
}

/* host order of cuda_src/rubiks.cpp:324-376: shift grad -> reduce -> normalise -> input grad */
void FN(oracle_shift3d_backward)(const REAL *x, const REAL *shift, const REAL *og, REAL *gin,
                                 REAL *gshift, int N, int T, int C, int H, int W, int To, int Ho,
                                 int Wo, int sT, int sH, int sW, int pT, int pH, int pW,
                                 int normalize_grad, REAL normalize_t_factor, int quantize) {
    FN(oracle_shift3d_backward_shift)(x, shift, og, gshift, N, T, C, H, W, To, Ho, Wo, sT, sH, sW,
                                      pT, pH, pW);
    if (normalize_grad) FN(oracle_normalize_shift_grad_3d)(gshift, C, normalize_t_factor);
    FN(oracle_shift3d_backward_input)(shift, og, gin, N, T, C, H, W, To, Ho, Wo, sT, sH, sW, pT, pH,
                                      pW, quantize);
}

/* ------------------------------------------------------------------ 2D ------------------------ */

/* cuda_src/rubiks2d_kernels.cu:69-73 */
static inline int FN(floor_fast)(REAL x) {
    int ix = (int)x;
    return ix - (x < ix);
}
/* cuda_src/rubiks2d_kernels.cu:76-82: round half away from zero */
static inline int FN(round_fast)(REAL x) {
    if (x < (REAL)0.0f) return (int)(x - (REAL)0.5f);
    return (int)(x + (REAL)0.5f);
}
static inline REAL FN(tap2d)(const REAL *p, int h, int w, int H, int W) {
    if (h < 0 || w < 0 || h >= H || w >= W) return (REAL)0;
    return p[(int64_t)h * W + w];
}
/* cuda_src/rubiks2d_kernels.cu:60-66 */
static inline REAL FN(interp2d)(REAL p00, REAL p01, REAL p10, REAL p11, REAL rh, REAL rw) {
    return p00 * (1 - rh) * (1 - rw) + p01 * (1 - rh) * rw + p10 * rh * (1 - rw) + p11 * rh * rw;
}

/*
 * cuda_src/rubiks2d_kernels.cu:94-145.  The quantize branch (:116-121) only writes in-bounds taps and
 * relies on the caller's pre-zeroed output (rubiksnet/utils.py:25-26); this restatement keeps that.
 */
void FN(oracle_shift2d_forward)(const REAL *x, const REAL *shift, REAL *out, int N, int C, int H,
                                int W, int Ho, int Wo, int sH, int sW, int pH, int pW,
                                int quantize) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
        for (int c = 0; c < C; ++c) {
            const REAL offh = shift[c], offw = shift[C + c];
            const REAL *xp = x + ((int64_t)n * C + c) * H * W;
            REAL *op = out + ((int64_t)n * C + c) * Ho * Wo;
            for (int ho = 0; ho < Ho; ++ho)
                for (int wo = 0; wo < Wo; ++wo) {
                    const int bh = ho * sH - pH, bw = wo * sW - pW;
                    if (quantize) {
                        const int th = FN(round_fast)(bh + offh), tw = FN(round_fast)(bw + offw);
                        if (th >= 0 && tw >= 0 && th < H && tw < W)
                            op[(int64_t)ho * Wo + wo] = xp[(int64_t)th * W + tw];
                        continue;
                    }
                    const int fh = FN(floor_fast)(offh), fw = FN(floor_fast)(offw);
                    const REAL rh = offh - fh, rw = offw - fw;
                    const int h0 = bh + fh, w0 = bw + fw;
                    op[(int64_t)ho * Wo + wo] = FN(interp2d)(
                        FN(tap2d)(xp, h0, w0, H, W), FN(tap2d)(xp, h0, w0 + 1, H, W),
                        FN(tap2d)(xp, h0 + 1, w0, H, W), FN(tap2d)(xp, h0 + 1, w0 + 1, H, W), rh, rw);
                }
        }
}

/*
 * cuda_src/rubiks2d_kernels.cu:147-266 + reduction rubiks.cpp:127-143: integer shifts (|r|<1e-7) use
 * 0.5 * central difference (:189-253).  Writes the UN-normalised [2,C] gradient.
 */
void FN(oracle_shift2d_backward_shift)(const REAL *x, const REAL *shift, const REAL *og,
                                       REAL *gshift, int N, int C, int H, int W, int Ho, int Wo,
                                       int sH, int sW, int pH, int pW) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < C; ++c) {
        const REAL offh = shift[c], offw = shift[C + c];
        const int fh = FN(floor_fast)(offh), fw = FN(floor_fast)(offw);
        REAL rh = offh - fh, rw = offw - fw;
        const REAL tol = (REAL)1e-7f;
        int ih = 0, iw = 0;
        if (tol > rh && rh > -tol) { ih = 1; rh = 0; }
        if (tol > rw && rw > -tol) { iw = 1; rw = 0; }
        double accH = 0, accW = 0;
        for (int n = 0; n < N; ++n) {
            const REAL *xp = x + ((int64_t)n * C + c) * H * W;
            const REAL *gp = og + ((int64_t)n * C + c) * Ho * Wo;
            for (int ho = 0; ho < Ho; ++ho)
                for (int wo = 0; wo < Wo; ++wo) {
                    const int h0 = ho * sH - pH + fh, w0 = wo * sW - pW + fw;
                    const REAL p00 = FN(tap2d)(xp, h0, w0, H, W), p01 = FN(tap2d)(xp, h0, w0 + 1, H, W);
                    const REAL p10 = FN(tap2d)(xp, h0 + 1, w0, H, W),
                               p11 = FN(tap2d)(xp, h0 + 1, w0 + 1, H, W);
                    REAL gH = (1 - rw) * (p10 - p00) + rw * (p11 - p01);
                    REAL gW = (1 - rh) * (p01 - p00) + rh * (p11 - p10);
                    if (ih || iw) {
                        /* P[a][b] = x(h0 + a - 1, w0 + b - 1); P[1][1] is the origin pixel */
#define P(a, b) FN(tap2d)(xp, h0 + (a)-1, w0 + (b)-1, H, W)
                        if (ih)
                            gH = (REAL)0.5f * ((1 - rw) * (P(2, 1) - P(0, 1)) + rw * (P(2, 2) - P(0, 2)));
                        if (iw)
                            gW = (REAL)0.5f * ((1 - rh) * (P(1, 2) - P(1, 0)) + rh * (P(2, 2) - P(2, 0)));
#undef P
                    }
                    const REAL up = gp[(int64_t)ho * Wo + wo];
                    accH += (double)(REAL)(gH * up);
                    accW += (double)(REAL)(gW * up);
                }
        }
        gshift[c] = (REAL)accH;
        gshift[C + c] = (REAL)accW;
    }
}

/* cuda_src/rubiks2d_kernels.cu:381-397 */
void FN(oracle_normalize_shift_grad_2d)(REAL *g, int C) {
    for (int c = 0; c < C; ++c) {
        const REAL gh = g[c], gw = g[C + c];
        const REAL mag = (REAL)sqrt((double)(gh * gh + gw * gw));
        if (mag > 0) {
            g[c] = gh / mag;
            g[C + c] = gw / mag;
        }
    }
}

static inline REAL FN(tap2d_adj)(const REAL *gp, int h, int w, int sH, int sW, int Ho, int Wo) {
    if (h % sH != 0 || w % sW != 0) return (REAL)0;
    h /= sH;
    w /= sW;
    if (h < 0 || w < 0 || h >= Ho || w >= Wo) return (REAL)0;
    return gp[(int64_t)h * Wo + w];
}

/*
 * cuda_src/rubiks2d_kernels.cu:269-379.  The quantize branch (:294-309) writes only in-bounds,
 * stride-aligned taps into the caller's pre-zeroed gin.
 */
void FN(oracle_shift2d_backward_input)(const REAL *shift, const REAL *og, REAL *gin, int N, int C,
                                       int H, int W, int Ho, int Wo, int sH, int sW, int pH, int pW,
                                       int quantize) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
        for (int c = 0; c < C; ++c) {
            const REAL sh = -shift[c], sw = -shift[C + c];
            const REAL *gp = og + ((int64_t)n * C + c) * Ho * Wo;
            REAL *ip = gin + ((int64_t)n * C + c) * H * W;
            for (int h = 0; h < H; ++h)
                for (int w = 0; w < W; ++w) {
                    const int bh = h + pH, bw = w + pW;
                    if (quantize) {
                        int th = FN(round_fast)(bh + sh), tw = FN(round_fast)(bw + sw);
                        if (th % sH == 0 && tw % sW == 0) {
                            th /= sH;
                            tw /= sW;
                            if (th >= 0 && tw >= 0 && th < Ho && tw < Wo)
                                ip[(int64_t)h * W + w] = gp[(int64_t)th * Wo + tw];
                        }
                        continue;
                    }
                    REAL v = 0;
                    if (sw == 0 && sh == 0) {
                        v = FN(tap2d_adj)(gp, bh, bw, sH, sW, Ho, Wo);
                    } else {
                        const int fh = FN(floor_fast)(sh), fw = FN(floor_fast)(sw);
                        const REAL rh = sh - fh, rw = sw - fw;
                        v = FN(interp2d)(FN(tap2d_adj)(gp, bh + fh, bw + fw, sH, sW, Ho, Wo),
                                         FN(tap2d_adj)(gp, bh + fh, bw + fw + 1, sH, sW, Ho, Wo),
                                         FN(tap2d_adj)(gp, bh + fh + 1, bw + fw, sH, sW, Ho, Wo),
                                         FN(tap2d_adj)(gp, bh + fh + 1, bw + fw + 1, sH, sW, Ho, Wo),
                                         rh, rw);
                    }
                    ip[(int64_t)h * W + w] = v;
                }
        }
}

/* host order of cuda_src/rubiks.cpp:94-155 */
void FN(oracle_shift2d_backward)(const REAL *x, const REAL *shift, const REAL *og, REAL *gin,
                                 REAL *gshift, int N, int C, int H, int W, int Ho, int Wo, int sH,
                                 int sW, int pH, int pW, int normalize_grad, int enable_shift_grad,
                                 int quantize) {
    if (enable_shift_grad) {
        FN(oracle_shift2d_backward_shift)(x, shift, og, gshift, N, C, H, W, Ho, Wo, sH, sW, pH, pW);
        if (normalize_grad) FN(oracle_normalize_shift_grad_2d)(gshift, C);
    }
    FN(oracle_shift2d_backward_input)(shift, og, gin, N, C, H, W, Ho, Wo, sH, sW, pH, pW, quantize);
}

#undef FN
#undef CAT
#undef CAT_
