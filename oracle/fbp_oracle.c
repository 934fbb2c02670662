/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference FBP convertor.
 *
 * Restates Recon/FBP_kernel.py (reference) in plain C so that parity tests and the
 * bench's cpu_baseline leg have a checker that finishes in seconds:
 *   fbp_oracle_weight      <- FBP.convert            FBP_kernel.py:99-105 (flip, D*cos(gamma), dtheta)
 *   fbp_oracle_ramp        <- conv_pj                FBP_kernel.py:125-131 (full convolve, "same" slice)
 *   fbp_oracle_backproject <- fbp_cpu                FBP_kernel.py:166-184 (fp64 trig, f32 accumulator)
 *   fbp_oracle_convert     <- FBP.convert            FBP_kernel.py:86-122
 * The tables (theta, nda, r, phi, h_RL, D*cos(nda)) are built by oracle/fbp_oracle.py with
 * the same numpy expressions as FBP.__init__ (FBP_kernel.py:27-84) and passed in.
 *
 * Threads split the PIXEL rows, never the views: each pixel accumulates its 2000 views in
 * ascending order exactly like the 1-thread reference, so the result does not depend on the
 * thread count (the reference's own prange over views races).
 *
 * Pinned against the reference run in the build container (oracle/make_golden.py ->
 * tests/golden/fbp_slice0.npz); deviation recorded in DESIGN.md.
 * Nothing under ipdm-pytorch_b200/ links or loads this file.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NV 2000
#define ND 912
#define NP 512

/* pj_out[v][n] = f32( f64( f32(pj_in[v][flip? ND-1-n : n] * wcos[n]) ) * dtheta ) */
void fbp_oracle_weight(const float *pj_in, float *pj_out, const float *wcos, double dtheta, int flip)
{
    for (int v = 0; v < NV; ++v)
        for (int n = 0; n < ND; ++n) {
            float a = pj_in[(size_t)v * ND + (flip ? ND - 1 - n : n)];
            float w = a * wcos[n];
            pj_out[(size_t)v * ND + n] = (float)((double)w * dtheta);
        }
}

/* q[v][n] = sum_m pj[v][m] * h[n - m + ND - 1], f32 accumulation in ascending m */
void fbp_oracle_ramp(const float *pj, float *q, const float *h)
{
#pragma omp parallel for schedule(static)
    for (int v = 0; v < NV; ++v) {
        const float *row = pj + (size_t)v * ND;
        for (int n = 0; n < ND; ++n) {
            float acc = 0.f;
            for (int m = 0; m < ND; ++m)
                acc += row[m] * h[n - m + ND - 1];
            q[(size_t)v * ND + n] = acc;
        }
    }
}

void fbp_oracle_backproject(const float *q, float *img, const double *theta, const float *nda,
                            const double *r, const double *phi, double D, double da)
{
    const double nda0 = (double)nda[0];
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < NP; ++i) {
        for (int j = 0; j < NP; ++j) {
            const double rr = r[i * NP + j], ph = phi[i * NP + j];
            float acc = 0.f;
            for (int t = 0; t < NV; ++t) {
                double beta = theta[t] - M_PI / 2;
                double th = M_PI / 2 + beta + ph;
                double sth = sin(th), cth = cos(th);
                double alpha = atan(rr * sth / (D + rr * cth));
                double u = (alpha - nda0) / da + 0.5;
                double curdet = floor(u);
                if (0 < curdet && curdet < ND) {
                    double lam = u - curdet;
                    double L = rr * sth / sin(alpha);
                    int k = (int)curdet;
                    const float *row = q + (size_t)t * ND;
                    acc = (float)((double)acc + ((1 - lam) * (double)row[k - 1] + lam * (double)row[k]) / (L * L));
                }
            }
            img[i * NP + j] = acc;
        }
    }
}

/* whole convertor for a batch; scratch is allocated here. returns 0 or -1 (allocation). */
int fbp_oracle_convert(const float *pj_in, float *img_out, int batch, int flip, const float *wcos, double dtheta,
                       const float *h, const double *theta, const float *nda, const double *r, const double *phi,
                       double D, double da)
{
    float *w = (float *)malloc(sizeof(float) * NV * ND);
    float *q = (float *)malloc(sizeof(float) * NV * ND);
    float *im = (float *)malloc(sizeof(float) * NP * NP);
    if (!w || !q || !im) { free(w); free(q); free(im); return -1; }
    for (int b = 0; b < batch; ++b) {
        fbp_oracle_weight(pj_in + (size_t)b * NV * ND, w, wcos, dtheta, flip);
        fbp_oracle_ramp(w, q, h);
        fbp_oracle_backproject(q, im, theta, nda, r, phi, D, da);
        float *dst = img_out + (size_t)b * NP * NP;
        for (int i = 0; i < NP; ++i)
            for (int j = 0; j < NP; ++j)
                dst[i * NP + j] = im[i * NP + (flip ? NP - 1 - j : j)];
    }
    free(w); free(q); free(im);
    return 0;
}
