/*
 * Plain-C CPU oracle for the anti-aliased activation path.  TEST INFRASTRUCTURE ONLY: built by
 * oracle/Makefile into oracle/_build/libafa_oracle.so and loaded only by tests/, smoke() and
 * bench.py's cpu_baseline leg.  The product (the CUDA library) never links or calls it.
 *
 * It restates the reference's algorithm in clamped-index closed form (a third, independent
 * formulation next to oracle/afa_oracle.py's op-by-op numpy and oracle/torch_path.py):
 *
 *   u[n] = 2 * sum_i  f_up[n + 5 - 2 i] * x[clamp(i, 0, T-1)]          0 <= n + 5 - 2 i <= 11
 *          (replicate pad 5/5 + conv_transpose1d stride 2 + x2 + crop 15/15,
 *           /root/reference/BigVGAN/alias_free_activation/resample.py:18-36)
 *   s[n] = u[n] + sin(alpha_eff * u[n])^2 / (beta_eff + 1e-9)
 *          (/root/reference/BigVGAN/activations.py:119-124; Snake: beta_eff = alpha_eff, :57-60)
 *   y[t] = sum_k f_dn[k] * s[clamp(2 t + k - 5, 0, 2T-1)]               k = 0..11
 *          (replicate pad 5/6 + conv1d stride 2,
 *           /root/reference/BigVGAN/alias_free_activation/filter.py:85-99)
 *
 * All arithmetic is double precision ("truth"); the *_f32io entry points only convert I/O.
 * Parity of this file is pinned by tests/test_oracle.py against tests/golden/ *.npz (outputs of the
 * unmodified reference, see tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define AFA_K 12
#define AFA_EPS 0.000000001

static inline long clampl(long v, long lo, long hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void eff_params(const double *alpha, const double *beta, int logscale, long c,
                       double *a_eff, double *b_eff) {
    double a = alpha[c];
    double b = beta ? beta[c] : alpha[c];
    if (logscale) { a = exp(a); b = exp(b); }
    *a_eff = a; *b_eff = b;
}

static void upsample_row(const double *x, long T, const double *f_up, double *u) {
    for (long n = 0; n < 2 * T; ++n) {
        double acc = 0.0;
        for (int k = 0; k < AFA_K; ++k) {
            long twice_i = n + 5 - k;               /* tap k touches x index (n + 5 - k) / 2 */
            if (twice_i & 1) continue;
            acc += f_up[k] * x[clampl(twice_i / 2, 0, T - 1)];   /* even => exact, also when negative */
        }
        u[n] = 2.0 * acc;
    }
}

int afa_oracle_fwd_f64(const double *x, double *y, const double *alpha, const double *beta,
                       int logscale, const double *f_up, const double *f_dn,
                       long B, long C, long T) {
    if (T <= 0 || B <= 0 || C <= 0) return 0;
    double *u = (double *)malloc(sizeof(double) * 2 * (size_t)T);
    if (!u) return -1;
    for (long r = 0; r < B * C; ++r) {
        const double *xr = x + r * T;
        double *yr = y + r * T;
        double a, b;
        eff_params(alpha, beta, logscale, r % C, &a, &b);
        const double ib = 1.0 / (b + AFA_EPS);
        upsample_row(xr, T, f_up, u);
        for (long n = 0; n < 2 * T; ++n) {
            double sn = sin(u[n] * a);
            u[n] = u[n] + ib * sn * sn;
        }
        for (long t = 0; t < T; ++t) {
            double acc = 0.0;
            for (int k = 0; k < AFA_K; ++k) acc += f_dn[k] * u[clampl(2 * t + k - 5, 0, 2 * T - 1)];
            yr[t] = acc;
        }
    }
    free(u);
    return 0;
}

/* gradients w.r.t. x and the RAW parameters (log-scale chain rule and Snake aliasing applied) */
int afa_oracle_bwd_f64(const double *x, const double *gy, double *gx, double *galpha, double *gbeta,
                       const double *alpha, const double *beta, int logscale,
                       const double *f_up, const double *f_dn, long B, long C, long T) {
    for (long c = 0; c < C; ++c) { galpha[c] = 0.0; if (gbeta) gbeta[c] = 0.0; }
    if (T <= 0 || B <= 0 || C <= 0) return 0;
    double *u = (double *)malloc(sizeof(double) * 2 * (size_t)T);
    double *gs = (double *)malloc(sizeof(double) * 2 * (size_t)T);
    if (!u || !gs) { free(u); free(gs); return -1; }
    for (long r = 0; r < B * C; ++r) {
        const long c = r % C;
        const double *xr = x + r * T, *gyr = gy + r * T;
        double *gxr = gx + r * T;
        double a, b;
        eff_params(alpha, beta, logscale, c, &a, &b);
        const double ib = 1.0 / (b + AFA_EPS);
        upsample_row(xr, T, f_up, u);
        memset(gs, 0, sizeof(double) * 2 * (size_t)T);
        for (long t = 0; t < T; ++t)
            for (int k = 0; k < AFA_K; ++k)
                gs[clampl(2 * t + k - 5, 0, 2 * T - 1)] += f_dn[k] * gyr[t];
        double ga = 0.0, gb = 0.0;
        for (long t = 0; t < T; ++t) gxr[t] = 0.0;
        for (long n = 0; n < 2 * T; ++n) {
            const double s2 = sin(2.0 * a * u[n]);
            const double s1 = sin(a * u[n]);
            ga += gs[n] * ib * u[n] * s2;
            gb -= gs[n] * s1 * s1 * ib * ib;
            const double gu = gs[n] * (1.0 + ib * a * s2);
            for (int k = 0; k < AFA_K; ++k) {
                long twice_i = n + 5 - k;
                if (twice_i & 1) continue;
                gxr[clampl(twice_i / 2, 0, T - 1)] += 2.0 * f_up[k] * gu;
            }
        }
        if (logscale) { ga *= a; gb *= b; }
        if (gbeta) { galpha[c] += ga; gbeta[c] += gb; }
        else { galpha[c] += ga + gb; }
    }
    free(u); free(gs);
    return 0;
}

/* float32 I/O wrappers (math stays double) */
int afa_oracle_fwd_f32io(const float *x, float *y, const float *alpha, const float *beta,
                         int logscale, const float *f_up, const float *f_dn,
                         long B, long C, long T) {
    size_t n = (size_t)B * C * T;
    double *xd = (double *)malloc(sizeof(double) * (n ? n : 1));
    double *yd = (double *)malloc(sizeof(double) * (n ? n : 1));
    double *ad = (double *)malloc(sizeof(double) * C);
    double *bd = beta ? (double *)malloc(sizeof(double) * C) : NULL;
    double fu[AFA_K], fd[AFA_K];
    if (!xd || !yd || !ad || (beta && !bd)) { free(xd); free(yd); free(ad); free(bd); return -1; }
    for (size_t i = 0; i < n; ++i) xd[i] = x[i];
    for (long c = 0; c < C; ++c) { ad[c] = alpha[c]; if (bd) bd[c] = beta[c]; }
    for (int k = 0; k < AFA_K; ++k) { fu[k] = f_up[k]; fd[k] = f_dn[k]; }
    int rc = afa_oracle_fwd_f64(xd, yd, ad, bd, logscale, fu, fd, B, C, T);
    for (size_t i = 0; i < n; ++i) y[i] = (float)yd[i];
    free(xd); free(yd); free(ad); free(bd);
    return rc;
}
