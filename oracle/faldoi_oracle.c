/* TEST INFRASTRUCTURE ONLY -- see faldoi_oracle.h.
 *
 * Plain-C restatement of the reference's global_faldoi solvers, written from
 * the reference's behaviour (file:line cited per function; paths relative to
 * /root/reference).  Data layout is ours (SoA planes, no PosNei /
 * DualVariables AoS); arithmetic keeps the reference's evaluation order and
 * float/double promotions so results are bit-comparable.  Build with
 * -ffp-contract=off (oracle/Makefile).  OpenMP is used only on loops whose
 * iterations are independent, so results do not depend on the thread count.
 */
#include "faldoi_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* optional per-iteration error trace of fo_tvl2 (tools/stripes_policy.py): "warp iteration err" lines */
static FILE *fo_err_trace = NULL;
void fo_set_err_trace(const char *path) {
    if (fo_err_trace) fclose(fo_err_trace);
    fo_err_trace = path ? fopen(path, "w") : NULL;
}

#define GRAD_IS_ZERO 1E-8 /* src/parameters.h:45 (a double literal: comparisons promote) */

static float *falloc(size_t n) {
    float *p = (float *)calloc(n ? n : 1, sizeof(float));
    if (!p) abort();
    return p;
}

/* ------------------------------------------------------------------ */
/* stencils: src/utils.cpp:239-283 (divergence), :285-323 (forward_gradient),
 * :367-423 (centered_gradient).  One branchy per-pixel form instead of the
 * reference's body/edge/corner loops; the fp32 association of every case is
 * the reference's.                                                      */
/* ------------------------------------------------------------------ */
void fo_forward_gradient(const float *f, float *fx, float *fy, int w, int h) {
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            const int p = j * w + i;
            fx[p] = (i < w - 1) ? f[p + 1] - f[p] : 0.f;
            fy[p] = (j < h - 1) ? f[p + w] - f[p] : 0.f;
        }
}

static inline float div_at(const float *a, const float *b, int i, int j, int w, int h) {
    const int p = j * w + i;
    const int first_c = (i == 0), last_c = (i == w - 1), first_r = (j == 0), last_r = (j == h - 1);
    if (!first_c && !last_c && !first_r && !last_r) return (a[p] - a[p - 1]) + (b[p] - b[p - w]);
    if (!first_c && !last_c) { /* first / last row, interior columns (:262-267) */
        if (first_r) return (a[p] - a[p - 1]) + b[p];
        return (a[p] - a[p - 1]) - b[p - w];
    }
    if (!first_r && !last_r) { /* first / last column, interior rows (:270-276) */
        if (first_c) return (a[p] + b[p]) - b[p - w];
        return (-a[p - 1] + b[p]) - b[p - w];
    }
    /* corners (:279-282) */
    if (first_c && first_r) return a[p] + b[p];
    if (last_c && first_r) return -a[p - 1] + b[p];
    if (first_c && last_r) return a[p] - b[p - w];
    return -a[p - 1] - b[p - w];
}

void fo_divergence(const float *a, const float *b, float *div, int w, int h) {
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) div[j * w + i] = div_at(a, b, i, j, w, h);
}

void fo_centered_gradient(const float *f, float *dx, float *dy, int w, int h) {
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            const int p = j * w + i;
            const int ir = (i < w - 1) ? p + 1 : p, il = (i > 0) ? p - 1 : p;
            const int jd = (j < h - 1) ? p + w : p, ju = (j > 0) ? p - w : p;
            dx[p] = (float)(0.5 * (f[ir] - f[il])); /* float difference, double product */
            dy[p] = (float)(0.5 * (f[jd] - f[ju]));
        }
}

/* ------------------------------------------------------------------ */
/* bicubic warp: src/bicubic_interpolation.c:103-111 (cell), :138-237 (at),
 * :245-266 (warp).  Neumann clamp; the row above uses sx (sic, :159).   */
/* ------------------------------------------------------------------ */
static inline float keys_cell(float v0, float v1, float v2, float v3, float t) {
    return (float)(v1 + 0.5 * t * (v2 - v0 + t * (2.0 * v0 - 5.0 * v1 + 4.0 * v2 - v3 + t * (3.0 * (v1 - v2) + v3 - v0))));
}

static inline int clampi(int x, int n, int *out) {
    if (x < 0) {
        *out = 1;
        return 0;
    }
    if (x >= n) {
        *out = 1;
        return n - 1;
    }
    return x;
}

static float bicubic_at(const float *img, float uu, float vv, int w, int h, int border_out) {
    const int sx = (uu < 0) ? -1 : 1, sy = (vv < 0) ? -1 : 1;
    const int xi = (int)uu, yi = (int)vv;
    int out = 0;
    const int x = clampi(xi, w, &out), y = clampi(yi, h, &out);
    const int mx = clampi(xi - sx, w, &out), my = clampi(yi - sx, h, &out);
    const int dx = clampi(xi + sx, w, &out), dy = clampi(yi + sy, h, &out);
    const int ddx = clampi(xi + 2 * sx, w, &out), ddy = clampi(yi + 2 * sy, h, &out);
    if (out && border_out) return 0.f;
    const float tx = uu - x, ty = vv - y;
    const int cols[4] = {mx, x, dx, ddx};
    float c[4];
    for (int k = 0; k < 4; k++)
        c[k] = keys_cell(img[cols[k] + w * my], img[cols[k] + w * y], img[cols[k] + w * dy], img[cols[k] + w * ddy], ty);
    return keys_cell(c[0], c[1], c[2], c[3], tx);
}

void fo_bicubic_warp(const float *img, const float *u, const float *v, float *out, int w, int h, int border_out) {
#pragma omp parallel for
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            const int p = j * w + i;
            const float uu = i + u[p], vv = j + v[p];
            out[p] = bicubic_at(img, uu, vv, w, h, border_out);
        }
}

/* ------------------------------------------------------------------ */
/* preprocessing                                                         */
/* ------------------------------------------------------------------ */
/* src/global_faldoi.cpp:1820-1827 */
void fo_rgb2gray(const float *rgb, int w, int h, float *out) {
    const int n = w * h;
    for (int i = 0; i < n; i++) out[i] = (float)(.299 * rgb[i] + .587 * rgb[n + i] + .114 * rgb[2 * n + i]);
}

static void minmax(const float *x, int n, float *mn, float *mx) {
    *mn = *mx = x[0];
    for (int i = 1; i < n; i++) {
        if (x[i] < *mn) *mn = x[i];
        if (x[i] > *mx) *mx = x[i];
    }
}

/* src/utils.cpp:743-781 as called at src/global_faldoi.cpp:2065, i.e. with
 * (I1,I2,I0) := (i0,i1,i_1): the joint "min" is the LARGER of min(i1) and
 * min(min(i_1),min(i0)) (:763).                                          */
void fo_normalize3(float *i0, float *i1, float *im1, int n) {
    float mn_m1, mx_m1, mn_0, mx_0, mn_1, mx_1;
    minmax(im1, n, &mn_m1, &mx_m1);
    minmax(i0, n, &mn_0, &mx_0);
    minmax(i1, n, &mn_1, &mx_1);
    const float max01 = (mx_m1 > mx_0) ? mx_m1 : mx_0;
    const float mx = (mx_1 > max01) ? mx_1 : max01;
    const float min01 = (mn_m1 < mn_0) ? mn_m1 : mn_0;
    const float mn = (mn_1 > min01) ? mn_1 : min01;
    const float den = mx - mn;
    if (den > 0)
        for (int i = 0; i < n; i++) {
            im1[i] = (im1[i] - mn) / den;
            i0[i] = (i0[i] - mn) / den;
            i1[i] = (i1[i] - mn) / den;
        }
}

/* src/utils.cpp:521-630: separable, in place, rows then columns; radius
 * size-1 with size=(int)(5*sigma)+1; "reflecting" pad that skips the edge
 * sample on the low side and repeats it on the high side (:569-573).     */
void fo_gaussian(float *img, int w, int h, float sigma) {
    const float den = 2 * sigma * sigma;
    const int size = (int)(5 * sigma) + 1;
    if (size > w || size > h) abort();
    float *B = falloc(size);
    for (int i = 0; i < size; i++) B[i] = (float)(1 / (sigma * sqrt(2.0 * 3.1415926)) * expf(-i * i / den));
    float norm = 0;
    for (int i = 0; i < size; i++) norm += B[i];
    norm *= 2;
    norm -= B[0];
    for (int i = 0; i < size; i++) B[i] /= norm;

    const int maxdim = (w > h) ? w : h;
    float *line = falloc((size_t)maxdim + 2 * size);
    for (int pass = 0; pass < 2; pass++) {
        const int len = pass ? h : w, cnt = pass ? w : h;
        const int stride = pass ? w : 1;
        for (int k = 0; k < cnt; k++) {
            float *base = img + (pass ? (size_t)k : (size_t)k * w);
            for (int i = 0; i < len; i++) line[size + i] = base[(size_t)i * stride];
            for (int i = 0; i < size; i++) {
                line[i] = base[(size_t)(size - i) * stride];
                line[size + len + i] = base[(size_t)(len - i - 1) * stride];
            }
            for (int i = size; i < size + len; i++) {
                float sum = B[0] * line[i];
                for (int j = 1; j < size; j++) sum += B[j] * (line[i - j] + line[i + j]);
                base[(size_t)(i - size) * stride] = sum;
            }
        }
    }
    free(line);
    free(B);
}

static inline float sq(float f) { return f * f; }

/* src/global_faldoi.cpp:906-932 */
void fo_image_to_lab(const float *rgb, int n, float *lab) {
    const float T = 0.008856;
    const float color_attenuation = 1.5f;
    for (int i = 0; i < n; i++) {
        const float r = rgb[i] / 255.f, g = rgb[i + n] / 255.f, b = rgb[i + 2 * n] / 255.f;
        float X = (float)(0.412453 * r + 0.357580 * g + 0.180423 * b);
        float Y = (float)(0.212671 * r + 0.715160 * g + 0.072169 * b);
        float Z = (float)(0.019334 * r + 0.119193 * g + 0.950227 * b);
        X = (float)(X / 0.950456);
        Z = (float)(Z / 1.088754);
        const float Y3 = (float)pow(Y, 1. / 3);
        const float fX = (float)(X > T ? pow(X, 1. / 3) : 7.787 * X + 16 / 116.);
        const float fY = (float)(Y > T ? (double)Y3 : 7.787 * Y + 16 / 116.);
        const float fZ = (float)(Z > T ? pow(Z, 1. / 3) : 7.787 * Z + 16 / 116.);
        const float L = (float)(Y > T ? 116 * Y3 - 16.0 : 903.3 * Y);
        const float A = 500 * (fX - fY);
        const float Bc = 200 * (fY - fZ);
        const float corr = expf(-color_attenuation * sq((float)(sq(L / 100) - 0.6)));
        lab[i] = L;
        lab[i + n] = A * corr;
        lab[i + 2 * n] = Bc * corr;
    }
}

/* src/global_faldoi.cpp:2049-2068 */
void fo_preprocess(const float *i0, const float *i1, const float *im1, int pd, int w, int h, float *i0n,
                   float *i1n, float *im1n) {
    const int n = w * h;
    if (pd != 1) {
        fo_rgb2gray(i0, w, h, i0n);
        fo_rgb2gray(i1, w, h, i1n);
        fo_rgb2gray(im1, w, h, im1n);
    } else {
        memcpy(i0n, i0, n * sizeof(float));
        memcpy(i1n, i1, n * sizeof(float));
        memcpy(im1n, im1, n * sizeof(float));
    }
    fo_normalize3(i0n, i1n, im1n, n);
    fo_gaussian(i0n, w, h, 0.90f);
    fo_gaussian(i1n, w, h, 0.90f);
    fo_gaussian(im1n, w, h, 0.90f);
}

/* src/parameters.h:20-31 via src/utils_preprocess.cpp:49-63 */
void fo_default_params(fo_params *p) {
    p->lambda = 40;
    p->theta = 0.3;
    p->tau = 0.125;
    p->beta = 0.025;
    p->alpha = 0.0706776435878;
    p->tau_u = 0.0739776273913;
    p->tau_eta = 0.0839911992024;
    p->tau_chi = 0.134077646787;
    p->mu = 1.4058686732;
    p->tol = 0.01;
}

/* ------------------------------------------------------------------ */
/* shared pieces of the primal-dual loops                                */
/* ------------------------------------------------------------------ */

/* warp I1 and its centred gradients with the current flow, then the
 * per-warp constants: src/global_faldoi.cpp:635-660 (same at :1225-1238) */
static void warp_and_constants(const float *I0, const float *I1, const float *I1x, const float *I1y,
                               const float *u1, const float *u2, float *I1w, float *I1wx, float *I1wy,
                               float *grad, float *rho_c, int w, int h) {
    const int n = w * h;
    fo_bicubic_warp(I1, u1, u2, I1w, w, h, 1);
    fo_bicubic_warp(I1x, u1, u2, I1wx, w, h, 1);
    fo_bicubic_warp(I1y, u1, u2, I1wy, w, h, 1);
    for (int i = 0; i < n; i++) {
        const float Ix2 = I1wx[i] * I1wx[i], Iy2 = I1wy[i] * I1wy[i];
        grad[i] = Ix2 + Iy2;
        rho_c[i] = I1w[i] - I1wx[i] * u1[i] - I1wy[i] * u2[i] - I0[i];
    }
}

/* thresholding operator TH: src/global_faldoi.cpp:690-718 (= :1254-1282) */
static void th_update(const float *u1, const float *u2, const float *rho_c, const float *grad, const float *Ix,
                      const float *Iy, float l_t, float *v1, float *v2, int n) {
#pragma omp parallel for
    for (int i = 0; i < n; i++) {
        const float rho = rho_c[i] + (Ix[i] * u1[i] + Iy[i] * u2[i]);
        float d1, d2;
        if (rho < -l_t * grad[i]) {
            d1 = l_t * Ix[i];
            d2 = l_t * Iy[i];
        } else if (rho > l_t * grad[i]) {
            d1 = -l_t * Ix[i];
            d2 = -l_t * Iy[i];
        } else if (grad[i] < GRAD_IS_ZERO) {
            d1 = d2 = 0;
        } else {
            const float fi = -rho / grad[i];
            d1 = fi * Ix[i];
            d2 = fi * Iy[i];
        }
        v1[i] = u1[i] + d1;
        v2[i] = u2[i] + d2;
    }
}

/* ------------------------------------------------------------------ */
/* TVL2 (methods 0,1): tvl2OF src/global_faldoi.cpp:556-882,
 * ofTVl2_getD :349-381, ofTVl2_getP :307-342                            */
/* ------------------------------------------------------------------ */
void fo_tvl2(const float *I0, const float *I1, float *u1, float *u2, float *xi11, float *xi12, float *xi21,
             float *xi22, float lambda, float theta, float tau, float tol, int w, int h, int warps,
             int max_iter, fo_log *log) {
    const int n = w * h;
    const float l_t = lambda * theta;
    float *buf = falloc((size_t)18 * n);
    float *I1x = buf, *I1y = buf + n, *I1w = buf + 2 * n, *I1wx = buf + 3 * n, *I1wy = buf + 4 * n;
    float *grad = buf + 5 * n, *rho_c = buf + 6 * n, *v1 = buf + 7 * n, *v2 = buf + 8 * n;
    float *ub1 = buf + 9 * n, *ub2 = buf + 10 * n, *g1x = buf + 11 * n, *g1y = buf + 12 * n;
    float *g2x = buf + 13 * n, *g2y = buf + 14 * n, *d1 = buf + 15 * n, *d2 = buf + 16 * n, *uN = buf + 17 * n;

    fo_centered_gradient(I1, I1x, I1y, w, h);
    for (int wp = 0; wp < warps; wp++) {
        warp_and_constants(I0, I1, I1x, I1y, u1, u2, I1w, I1wx, I1wy, grad, rho_c, w, h);
        memcpy(ub1, u1, n * sizeof(float));
        memcpy(ub2, u2, n * sizeof(float));
        int it = 0;
        float err = INFINITY;
        while (err > tol * tol && it < max_iter) {
            it++;
            th_update(u1, u2, rho_c, grad, I1wx, I1wy, l_t, v1, v2, n);
            fo_forward_gradient(ub1, g1x, g1y, w, h);
            fo_forward_gradient(ub2, g2x, g2y, w, h);
#pragma omp parallel for
            for (int i = 0; i < n; i++) { /* dual step, normalised by the OLD |xi| (:364-378) */
                const float a = xi11[i] * xi11[i], b = xi12[i] * xi12[i], c = xi21[i] * xi21[i], d = xi22[i] * xi22[i];
                float nrm = sqrtf(a + b + c + d);
                nrm = (1 > nrm) ? 1 : nrm;
                xi11[i] = (xi11[i] + tau * g1x[i]) / nrm;
                xi12[i] = (xi12[i] + tau * g1y[i]) / nrm;
                xi21[i] = (xi21[i] + tau * g2x[i]) / nrm;
                xi22[i] = (xi22[i] + tau * g2y[i]) / nrm;
            }
            fo_divergence(xi11, xi12, d1, w, h);
            fo_divergence(xi21, xi22, d2, w, h);
#pragma omp parallel for
            for (int i = 0; i < n; i++) { /* primal step + extrapolation (:325-335, :780-783) */
                const float u1k = u1[i], u2k = u2[i];
                u1[i] = u1k - tau * (-d1[i] + (u1k - v1[i]) / theta);
                u2[i] = u2k - tau * (-d2[i] + (u2k - v2[i]) / theta);
                uN[i] = (u1[i] - u1k) * (u1[i] - u1k) + (u2[i] - u2k) * (u2[i] - u2k);
                ub1[i] = 2 * u1[i] - u1k;
                ub2[i] = 2 * u2[i] - u2k;
            }
            float mx = uN[0]; /* err is the MAX of |du|^2 (getminmax, :338-341) */
            for (int i = 1; i < n; i++)
                if (uN[i] > mx) mx = uN[i];
            err = mx;
            if (fo_err_trace) fprintf(fo_err_trace, "%d %d %.9g\n", wp, it, err);
        }
        if (log && wp < FO_MAX_WARPS) {
            log->iters[wp] = it;
            log->err[wp] = err;
        }
    }
    free(buf);
}

/* ------------------------------------------------------------------ */
/* CSAD data term: rank selection of src/global_faldoi.cpp:1549-1570
 * (hyp=1: grad = hypot(Ix^2+Iy^2, 0.01), TV-CSAD) and :1734-1758
 * (hyp=0: scale = sqrt(Ix^2+Iy^2) guarded by grad > 1e-8, NLTV-CSAD).
 * Neighbour enumeration of initialize_pos_nei :1332-1374: rows k=-3..3
 * outer, columns l=-3..3 inner, centre skipped, only in-image ones kept.
 * bnei is [48][n] (slot-major), filled per warp.                         */
/* ------------------------------------------------------------------ */
#define CSAD_R 3
#define CSAD_N 48

static int cmp_float(const void *a, const void *b) {
    const float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* per-warp: scale[] and b_j (:1514-1534 / :1698-1723) */
static void csad_constants(const float *I0, const float *I1w, const float *Ix, const float *Iy, const float *u1,
                           const float *u2, int hyp, float *grad, float *scale, float *bnei, int w, int h) {
    const int n = w * h;
#pragma omp parallel for
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            const int p = j * w + i;
            const float Ix2 = Ix[p] * Ix[p], Iy2 = Iy[p] * Iy[p];
            if (hyp) {
                grad[p] = (float)hypot(Ix2 + Iy2, 0.01);
                scale[p] = grad[p];
            } else {
                grad[p] = Ix2 + Iy2;
                if (!(grad[p] > GRAD_IS_ZERO)) continue; /* b keeps its previous content, never read */
                scale[p] = sqrtf(grad[p]);
            }
            int s = 0;
            for (int k = -CSAD_R; k <= CSAD_R; k++)
                for (int l = -CSAD_R; l <= CSAD_R; l++) {
                    if (k == 0 && l == 0) continue;
                    const int r = j + k, c = i + l;
                    if (c >= 0 && c < w && r >= 0 && r < h) {
                        const int q = r * w + c;
                        bnei[(size_t)s * n + p] =
                            (I0[p] - I0[q] - I1w[p] + I1w[q] + Ix[p] * u1[p] + Iy[p] * u2[p]) / scale[p];
                    }
                    s++;
                }
        }
}

static void csad_v_update(const float *u1, const float *u2, const float *Ix, const float *Iy, const float *grad,
                          const float *scale, const float *bnei, int hyp, float l_t, float *v1, float *v2, int w,
                          int h) {
    const int n = w * h;
#pragma omp parallel for
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            const int p = j * w + i;
            v1[p] = u1[p];
            v2[p] = u2[p];
            if (!hyp && !(grad[p] > GRAD_IS_ZERO)) continue;
            float cand[2 * CSAD_N + 1];
            int it = 0, s = 0, np = 0;
            for (int k = -CSAD_R; k <= CSAD_R; k++)
                for (int l = -CSAD_R; l <= CSAD_R; l++) {
                    if (k == 0 && l == 0) continue;
                    const int r = j + k, c = i + l;
                    if (c >= 0 && c < w && r >= 0 && r < h) {
                        cand[it++] = -(bnei[(size_t)s * n + p] - (Ix[p] * u1[p] + Iy[p] * u2[p]) / scale[p]);
                        np++;
                    }
                    s++;
                }
            for (int m = 0; m < np + 1; m++) cand[it++] = (np - 2 * m) * l_t * scale[p];
            qsort(cand, it, sizeof(float), cmp_float);
            const float med = cand[it / 2 + 1]; /* one past the median (upstream TODO :1567) */
            v1[p] = u1[p] - Ix[p] * med / scale[p];
            v2[p] = u2[p] - Iy[p] * med / scale[p];
        }
}

/* ------------------------------------------------------------------ */
/* TV-CSAD (methods 4,5): tvcsad_PD src/global_faldoi.cpp:1449-1637,
 * tvcsad_getD :1428-1446 (row-wise projection), tvcsad_getP :1386-1421
 * (err = mean |du|^2; the reference accumulates it with a racy +=, the
 * oracle defines it as the in-order fp32 sum)                           */
/* ------------------------------------------------------------------ */
void fo_tvcsad(const float *I0, const float *I1, float *u1, float *u2, float *xi11, float *xi12, float *xi21,
               float *xi22, float lambda, float theta, float tau, float tol, int w, int h, int warps,
               int max_iter, fo_log *log) {
    const int n = w * h;
    const float l_t = lambda * theta;
    float *buf = falloc((size_t)18 * n);
    float *bnei = falloc((size_t)CSAD_N * n);
    float *I1x = buf, *I1y = buf + n, *I1w = buf + 2 * n, *I1wx = buf + 3 * n, *I1wy = buf + 4 * n;
    float *grad = buf + 5 * n, *scale = buf + 6 * n, *v1 = buf + 7 * n, *v2 = buf + 8 * n;
    float *ub1 = buf + 9 * n, *ub2 = buf + 10 * n, *g1x = buf + 11 * n, *g1y = buf + 12 * n;
    float *g2x = buf + 13 * n, *g2y = buf + 14 * n, *d1 = buf + 15 * n, *d2 = buf + 16 * n;

    fo_centered_gradient(I1, I1x, I1y, w, h);
    for (int wp = 0; wp < warps; wp++) {
        fo_bicubic_warp(I1, u1, u2, I1w, w, h, 1);
        fo_bicubic_warp(I1x, u1, u2, I1wx, w, h, 1);
        fo_bicubic_warp(I1y, u1, u2, I1wy, w, h, 1);
        csad_constants(I0, I1w, I1wx, I1wy, u1, u2, 1, grad, scale, bnei, w, h);
        memcpy(ub1, u1, n * sizeof(float));
        memcpy(ub2, u2, n * sizeof(float));
        int it = 0;
        float err = INFINITY;
        while (err > tol * tol && it < max_iter) {
            it++;
            csad_v_update(u1, u2, I1wx, I1wy, grad, scale, bnei, 1, l_t, v1, v2, w, h);
            fo_forward_gradient(ub1, g1x, g1y, w, h);
            fo_forward_gradient(ub2, g2x, g2y, w, h);
#pragma omp parallel for
            for (int i = 0; i < n; i++) {
                float n1 = hypotf(xi11[i], xi12[i]), n2 = hypotf(xi21[i], xi22[i]);
                n1 = (1 > n1) ? 1 : n1;
                n2 = (1 > n2) ? 1 : n2;
                xi11[i] = (xi11[i] + tau * g1x[i]) / n1;
                xi12[i] = (xi12[i] + tau * g1y[i]) / n1;
                xi21[i] = (xi21[i] + tau * g2x[i]) / n2;
                xi22[i] = (xi22[i] + tau * g2y[i]) / n2;
            }
            fo_divergence(xi11, xi12, d1, w, h);
            fo_divergence(xi21, xi22, d2, w, h);
            float acc = 0.f;
            for (int i = 0; i < n; i++) {
                const float u1k = u1[i], u2k = u2[i];
                u1[i] = u1k - tau * (-d1[i] + (u1k - v1[i]) / theta);
                u2[i] = u2k - tau * (-d2[i] + (u2k - v2[i]) / theta);
                acc += (u1[i] - u1k) * (u1[i] - u1k) + (u2[i] - u2k) * (u2[i] - u2k);
                ub1[i] = 2 * u1[i] - u1k;
                ub2[i] = 2 * u2[i] - u2k;
            }
            err = acc / n;
        }
        if (log && wp < FO_MAX_WARPS) {
            log->iters[wp] = it;
            log->err[wp] = err;
        }
    }
    free(bnei);
    free(buf);
}

/* ------------------------------------------------------------------ */
/* NLTV regulariser.  Slots s=0..23 <-> offsets (k,l), rows k=-2..2 outer,
 * columns l=-2..2 inner, centre skipped, slot counter advances for
 * out-of-image offsets too; reciprocal slot of s is 23-s
 * (initialize_dual_variables src/global_faldoi.cpp:996-1054).
 * Weights: get_weight_2 :980-994 with sigma_colour=5 (:954-978),
 * sigma_spatial=2 (:943-952).  wgt is [24][n], wt is [n], P/Q are [24][n].*/
/* ------------------------------------------------------------------ */
#define NL_R 2
#define NL_N 24

static void nltv_init(const float *lab, int pd, int w, int h, float *wgt, float *wt) {
    const int n = w * h;
#pragma omp parallel for
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            const int p = j * w + i;
            int s = 0;
            float ne = 0.0;
            for (int k = -NL_R; k <= NL_R; k++)
                for (int l = -NL_R; l <= NL_R; l++) {
                    if (k == 0 && l == 0) continue;
                    const int r = j + k, c = i + l;
                    if (c >= 0 && c < w && r >= 0 && r < h) {
                        float difI = 0.0;
                        for (int m = 0; m < pd; m++) {
                            const float aux = lab[(size_t)m * n + p] - lab[(size_t)m * n + r * w + c];
                            difI += aux * aux;
                        }
                        difI = sqrtf(difI);
                        const float wc = expf(-difI / 5.f);
                        const float difS = (float)hypot((double)l, (double)k);
                        const float ws = expf(-difS / 2.f);
                        const float wv = sqrtf(wc * ws);
                        wgt[(size_t)s * n + p] = wv;
                        ne += wv;
                    } else {
                        wgt[(size_t)s * n + p] = -2.0f; /* marks an out-of-image slot */
                    }
                    s++;
                }
            wt[p] = ne;
        }
}

/* ofnltv_getD :1127-1174, non_local_divergence :1056-1079, ofnltv_getP :1090-1120 */
static float nltv_step(float *u1, float *u2, float *ub1, float *ub2, const float *v1, const float *v2,
                       float *P, float *Q, const float *wgt, const float *wt, float *dP, float *dQ, float theta,
                       float tau, int w, int h) {
    const int n = w * h;
#pragma omp parallel for
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            const int p = j * w + i;
            int s = 0;
            for (int k = -NL_R; k <= NL_R; k++)
                for (int l = -NL_R; l <= NL_R; l++) {
                    if (k == 0 && l == 0) continue;
                    const int r = j + k, c = i + l;
                    if (c >= 0 && c < w && r >= 0 && r < h) {
                        const int q = r * w + c;
                        const size_t sp = (size_t)s * n + p;
                        const float g1 = wgt[sp] * (ub1[p] - ub1[q]) / wt[p];
                        P[sp] = (P[sp] + tau * g1) / (1 + tau * sqrtf(g1 * g1));
                        const float g2 = wgt[sp] * (ub2[p] - ub2[q]) / wt[p];
                        Q[sp] = (Q[sp] + tau * g2) / (1 + tau * sqrtf(g2 * g2));
                    }
                    s++;
                }
        }
#pragma omp parallel for
    for (int j = 0; j < h; j++)
        for (int i = 0; i < w; i++) {
            const int p = j * w + i;
            int s = 0;
            float a = 0.0, b = 0.0;
            for (int k = -NL_R; k <= NL_R; k++)
                for (int l = -NL_R; l <= NL_R; l++) {
                    if (k == 0 && l == 0) continue;
                    const int r = j + k, c = i + l;
                    if (c >= 0 && c < w && r >= 0 && r < h) {
                        const int q = r * w + c;
                        const size_t sp = (size_t)s * n + p, rq = (size_t)(NL_N - 1 - s) * n + q;
                        a += wgt[sp] * (P[sp] - P[rq]);
                        b += wgt[sp] * (Q[sp] - Q[rq]);
                    }
                    s++;
                }
            dP[p] = a / wt[p];
            dQ[p] = b / wt[p];
        }
    float acc = 0.f;
    for (int i = 0; i < n; i++) {
        const float u1k = u1[i], u2k = u2[i];
        u1[i] = u1k - tau * (dP[i] + (u1k - v1[i]) / theta); /* +div: sign differs from TV (:1110) */
        u2[i] = u2k - tau * (dQ[i] + (u2k - v2[i]) / theta);
        acc += (u1[i] - u1k) * (u1[i] - u1k) + (u2[i] - u2k) * (u2[i] - u2k);
        ub1[i] = 2 * u1[i] - u1k;
        ub2[i] = 2 * u2[i] - u2k;
    }
    return acc / n;
}

/* NLTV-L1 (methods 2,3): nltvl1_PD src/global_faldoi.cpp:1177-1328; fixed max_iter iterations (:1249) */
void fo_nltv(const float *I0, const float *I1, const float *lab, float *u1, float *u2, float lambda,
             float theta, float tau, int w, int h, int warps, int max_iter, fo_log *log) {
    const int n = w * h;
    const float l_t = lambda * theta;
    float *buf = falloc((size_t)14 * n);
    float *wgt = falloc((size_t)NL_N * n), *P = falloc((size_t)NL_N * n), *Q = falloc((size_t)NL_N * n);
    float *I1x = buf, *I1y = buf + n, *I1w = buf + 2 * n, *I1wx = buf + 3 * n, *I1wy = buf + 4 * n;
    float *grad = buf + 5 * n, *rho_c = buf + 6 * n, *v1 = buf + 7 * n, *v2 = buf + 8 * n;
    float *ub1 = buf + 9 * n, *ub2 = buf + 10 * n, *dP = buf + 11 * n, *dQ = buf + 12 * n, *wt = buf + 13 * n;

    nltv_init(lab, 3, w, h, wgt, wt);
    fo_centered_gradient(I1, I1x, I1y, w, h);
    for (int wp = 0; wp < warps; wp++) {
        warp_and_constants(I0, I1, I1x, I1y, u1, u2, I1w, I1wx, I1wy, grad, rho_c, w, h);
        memcpy(ub1, u1, n * sizeof(float));
        memcpy(ub2, u2, n * sizeof(float));
        int it = 0;
        float err = INFINITY;
        while (it < max_iter) {
            it++;
            th_update(u1, u2, rho_c, grad, I1wx, I1wy, l_t, v1, v2, n);
            err = nltv_step(u1, u2, ub1, ub2, v1, v2, P, Q, wgt, wt, dP, dQ, theta, tau, w, h);
        }
        if (log && wp < FO_MAX_WARPS) {
            log->iters[wp] = it;
            log->err[wp] = err;
        }
    }
    free(Q);
    free(P);
    free(wgt);
    free(buf);
}

/* NLTV-CSAD (methods 6,7): nltvcsad_PD src/global_faldoi.cpp:1642-1808 */
void fo_nltvcsad(const float *I0, const float *I1, const float *lab, float *u1, float *u2, float lambda,
                 float theta, float tau, int w, int h, int warps, int max_iter, fo_log *log) {
    const int n = w * h;
    const float l_t = lambda * theta;
    float *buf = falloc((size_t)14 * n);
    float *wgt = falloc((size_t)NL_N * n), *P = falloc((size_t)NL_N * n), *Q = falloc((size_t)NL_N * n);
    float *bnei = falloc((size_t)CSAD_N * n);
    float *I1x = buf, *I1y = buf + n, *I1w = buf + 2 * n, *I1wx = buf + 3 * n, *I1wy = buf + 4 * n;
    float *grad = buf + 5 * n, *scale = buf + 6 * n, *v1 = buf + 7 * n, *v2 = buf + 8 * n;
    float *ub1 = buf + 9 * n, *ub2 = buf + 10 * n, *dP = buf + 11 * n, *dQ = buf + 12 * n, *wt = buf + 13 * n;

    nltv_init(lab, 3, w, h, wgt, wt);
    fo_centered_gradient(I1, I1x, I1y, w, h);
    for (int wp = 0; wp < warps; wp++) {
        fo_bicubic_warp(I1, u1, u2, I1w, w, h, 1);
        fo_bicubic_warp(I1x, u1, u2, I1wx, w, h, 1);
        fo_bicubic_warp(I1y, u1, u2, I1wy, w, h, 1);
        csad_constants(I0, I1w, I1wx, I1wy, u1, u2, 0, grad, scale, bnei, w, h);
        memcpy(ub1, u1, n * sizeof(float));
        memcpy(ub2, u2, n * sizeof(float));
        int it = 0;
        float err = INFINITY;
        while (it < max_iter) {
            it++;
            csad_v_update(u1, u2, I1wx, I1wy, grad, scale, bnei, 0, l_t, v1, v2, w, h);
            err = nltv_step(u1, u2, ub1, ub2, v1, v2, P, Q, wgt, wt, dP, dQ, theta, tau, w, h);
        }
        if (log && wp < FO_MAX_WARPS) {
            log->iters[wp] = it;
            log->err[wp] = err;
        }
    }
    free(bnei);
    free(Q);
    free(P);
    free(wgt);
    free(buf);
}

/* ------------------------------------------------------------------ */
/* TVL2 + occlusions (method 8): guided_tvl2coupled_occ
 * src/tvl2_model_occ.cpp:492-779, tvl2coupled_get_xi_patch :312-407,
 * tvl2coupled_get_chi_patch :411-484, init_weight src/utils.cpp:838-852,
 * with the patch = whole image (src/global_faldoi.cpp:2093-2098).
 * TARGET DEFINITION: eta1 = eta2 = 0 and div_u = 0 at entry (the reference
 * leaves them uninitialised on this path).          */
/* ------------------------------------------------------------------ */
#define OCC_ITER_XI 25
#define OCC_ITER_CHI 25
#define OCC_THRESHOLD 0.6

void fo_tvl2occ(const float *I0, const float *I1, const float *Im1, float *u1, float *u2, float *chi,
                const fo_params *pr, int w, int h, int warps, int max_iter, fo_log *log) {
    const int n = w * h;
    const float alpha = pr->alpha, theta = pr->theta, lambda = pr->lambda, beta = pr->beta;
    const float l_t = lambda * theta;
    const float tau_theta = pr->tau_u / theta;
    enum { NB = 44 };
    float *buf = falloc((size_t)NB * n);
    int b = 0;
#define PL() (buf + (size_t)(b++) * n)
    float *xi11 = PL(), *xi12 = PL(), *xi21 = PL(), *xi22 = PL(), *uba1 = PL(), *uba2 = PL();
    float *I1x = PL(), *I1y = PL(), *Jx = PL(), *Jy = PL(), *I0x = PL(), *I0y = PL(), *g = PL();
    float *I1w = PL(), *I1wx = PL(), *I1wy = PL(), *Jw = PL(), *Jwx = PL(), *Jwy = PL();
    float *grad_1 = PL(), *grad__1 = PL(), *rho_c1 = PL(), *rho_c_1 = PL(), *v1 = PL(), *v2 = PL();
    float *chix = PL(), *chiy = PL(), *gx11 = PL(), *gx12 = PL(), *gx21 = PL(), *gx22 = PL();
    float *dv1 = PL(), *dv2 = PL(), *vi1 = PL(), *vi2 = PL(), *a1 = PL(), *b1 = PL(), *a2 = PL(), *b2 = PL();
    float *F = PL(), *G = PL(), *eta1 = PL(), *eta2 = PL(), *dN = PL();
#undef PL
    float *ge1 = gx11, *ge2 = gx12, *dge = dv1; /* scratch reuse: not live at the same time */
    /* div_u == 0 (target definition): beta*div_u is the float product beta*0 = 0 */

    for (int i = 0; i < n; i++) { /* :592-601 */
        uba1[i] = -u1[i];
        uba2[i] = -u2[i];
    }
    fo_centered_gradient(I1, I1x, I1y, w, h);
    fo_centered_gradient(Im1, Jx, Jy, w, h);
    fo_centered_gradient(I0, I0x, I0y, w, h);
    for (int i = 0; i < n; i++) { /* init_weight, gamma = 0.05 */
        const float gr = sqrtf(I0x[i] * I0x[i] + I0y[i] * I0y[i]);
        g[i] = 1 / (1 + 0.05f * gr);
    }

    for (int wp = 0; wp < warps; wp++) {
        fo_bicubic_warp(I1, u1, u2, I1w, w, h, 0);
        fo_bicubic_warp(I1x, u1, u2, I1wx, w, h, 0);
        fo_bicubic_warp(I1y, u1, u2, I1wy, w, h, 0);
        fo_bicubic_warp(Im1, uba1, uba2, Jw, w, h, 0);
        fo_bicubic_warp(Jx, uba1, uba2, Jwx, w, h, 0);
        fo_bicubic_warp(Jy, uba1, uba2, Jwy, w, h, 0);
        for (int i = 0; i < n; i++) { /* :628-648 */
            grad_1[i] = I1wx[i] * I1wx[i] + I1wy[i] * I1wy[i];
            grad__1[i] = Jwx[i] * Jwx[i] + Jwy[i] * Jwy[i];
            rho_c1[i] = I1w[i] - I1wx[i] * u1[i] - I1wy[i] * u2[i] - I0[i];
            rho_c_1[i] = Jw[i] - Jwx[i] * u1[i] - Jwy[i] * u2[i] - I0[i];
        }
        int it = 0;
        float err = INFINITY;
        while (err > pr->tol * pr->tol && it < max_iter) {
            it++;
#pragma omp parallel for
            for (int i = 0; i < n; i++) { /* v update with the chi switch (:657-713) */
                const float rho_1 = rho_c1[i] + I1wx[i] * u1[i] + I1wy[i] * u2[i];
                const float rho__1 = rho_c_1[i] + Jwx[i] * u1[i] + Jwy[i] * u2[i];
                int eps;
                float alpha_i, mu, Lambda, grad, Iwx, Iwy, rho;
                if (chi[i] == 0) {
                    eps = 1;
                    alpha_i = 1;
                    mu = l_t;
                    Lambda = rho_1;
                    grad = grad_1[i];
                    Iwx = I1wx[i];
                    Iwy = I1wy[i];
                    rho = rho_1;
                } else {
                    eps = -1;
                    alpha_i = 1 / (1 + alpha * theta);
                    mu = l_t / (1 + alpha * theta);
                    Lambda = rho__1 + alpha * theta / (1 + alpha * theta) * (u1[i] * Jwx[i] + u2[i] * Jwy[i]);
                    grad = grad__1[i];
                    Iwx = Jwx[i];
                    Iwy = Jwy[i];
                    rho = rho__1;
                }
                if (Lambda > mu * grad) {
                    v1[i] = alpha_i * u1[i] - mu * eps * Iwx;
                    v2[i] = alpha_i * u2[i] - mu * eps * Iwy;
                } else if (Lambda < -mu * grad) {
                    v1[i] = alpha_i * u1[i] + mu * eps * Iwx;
                    v2[i] = alpha_i * u2[i] + mu * eps * Iwy;
                } else if (grad < GRAD_IS_ZERO) {
                    v1[i] = u1[i];
                    v2[i] = u2[i];
                } else {
                    v1[i] = u1[i] - eps * rho * Iwx / grad;
                    v2[i] = u2[i] - eps * rho * Iwy / grad;
                }
            }
            fo_forward_gradient(chi, chix, chiy, w, h);
            /* ITER_XI-1 Chambolle sweeps + final divergence (:340-406) */
            for (int k = 1; k <= OCC_ITER_XI; k++) {
                for (int i = 0; i < n; i++) {
                    gx11[i] = g[i] * xi11[i];
                    gx12[i] = g[i] * xi12[i];
                    gx21[i] = g[i] * xi21[i];
                    gx22[i] = g[i] * xi22[i];
                }
                fo_divergence(gx11, gx12, dv1, w, h);
                fo_divergence(gx21, gx22, dv2, w, h);
                if (k == OCC_ITER_XI) break;
                for (int i = 0; i < n; i++) {
                    vi1[i] = v1[i] + theta * dv1[i] + theta * beta * chix[i];
                    vi2[i] = v2[i] + theta * dv2[i] + theta * beta * chiy[i];
                }
                fo_forward_gradient(vi1, a1, b1, w, h);
                fo_forward_gradient(vi2, a2, b2, w, h);
#pragma omp parallel for
                for (int i = 0; i < n; i++) {
                    const float e11 = g[i] * a1[i], e12 = g[i] * b1[i];
                    const float n1 = sqrtf(e11 * e11 + e12 * e12);
                    xi11[i] = (xi11[i] + tau_theta * e11) / (1 + tau_theta * n1);
                    xi12[i] = (xi12[i] + tau_theta * e12) / (1 + tau_theta * n1);
                    const float e21 = g[i] * a2[i], e22 = g[i] * b2[i];
                    const float n2 = sqrtf(e21 * e21 + e22 * e22);
                    xi21[i] = (xi21[i] + tau_theta * e21) / (1 + tau_theta * n2);
                    xi22[i] = (xi22[i] + tau_theta * e22) / (1 + tau_theta * n2);
                }
            }
            for (int i = 0; i < n; i++) { /* u update, F, G (:726-751) */
                const float u1k = u1[i], u2k = u2[i];
                u1[i] = v1[i] + theta * dv1[i] + theta * beta * chix[i];
                u2[i] = v2[i] + theta * dv2[i] + theta * beta * chiy[i];
                dN[i] = (u1[i] - u1k) * (u1[i] - u1k) + (u2[i] - u2k) * (u2[i] - u2k);
                const float rho__1 = rho_c_1[i] + Jwx[i] * v1[i] + Jwy[i] * v2[i];
                const float rho_1 = rho_c1[i] + I1wx[i] * v1[i] + I1wy[i] * v2[i];
                F[i] = lambda * (fabsf(rho__1) - fabsf(rho_1));
                G[i] = alpha / 2 * (v1[i] * v1[i] + v2[i] * v2[i]);
            }
            /* ITER_CHI-1 primal-dual sweeps for chi, then hard threshold (:431-483) */
            for (int k = 1; k < OCC_ITER_CHI; k++) {
#pragma omp parallel for
                for (int i = 0; i < n; i++) {
                    const float e1 = eta1[i] + pr->mu * pr->tau_eta * g[i] * chix[i];
                    const float e2 = eta2[i] + pr->mu * pr->tau_eta * g[i] * chiy[i];
                    const float ne = sqrtf(e1 * e1 + e2 * e2);
                    if (ne <= 1) {
                        eta1[i] = e1;
                        eta2[i] = e2;
                    } else {
                        eta1[i] = e1 / ne;
                        eta2[i] = e2 / ne;
                    }
                    ge1[i] = g[i] * eta1[i];
                    ge2[i] = g[i] * eta2[i];
                }
                fo_divergence(ge1, ge2, dge, w, h);
                for (int i = 0; i < n; i++) {
                    const float div_u = 0.f;
                    const float c = chi[i] + pr->tau_chi * (pr->mu * dge[i] - beta * div_u - F[i] - G[i]);
                    const float lo = (c < 1) ? c : 1;
                    chi[i] = (lo > 0) ? lo : 0;
                }
                fo_forward_gradient(chi, chix, chiy, w, h);
            }
            for (int i = 0; i < n; i++) chi[i] = (chi[i] > OCC_THRESHOLD) ? 1 : 0;
            float mx = 0;
            for (int i = 0; i < n; i++)
                if (mx < dN[i]) mx = dN[i];
            err = mx;
        }
        if (log && wp < FO_MAX_WARPS) {
            log->iters[wp] = it;
            log->err[wp] = err;
        }
    }
    free(buf);
}

/* method switch + hard-coded constants of main(): src/global_faldoi.cpp:2132-2167.
 * Methods 0-7 always use 400 iterations (MAX_ITERATIONS_GLOBAL), only method 8
 * honours glb_iters (src/tvl2_model_occ.cpp:653).                        */
int fo_global_solve(int method, const float *I0, const float *I1, const float *Im1, const float *lab,
                    float *u, float *chi, const fo_params *p, int w, int h, int warps, int glb_iters,
                    fo_log *log) {
    const int n = w * h;
    float *u1 = u, *u2 = u + n;
    if (method == 0 || method == 1 || method == 4 || method == 5) {
        float *xi = falloc((size_t)4 * n);
        if (method <= 1)
            fo_tvl2(I0, I1, u1, u2, xi, xi + n, xi + 2 * n, xi + 3 * n, p->lambda, p->theta, p->tau, p->tol, w, h,
                    warps, 400, log);
        else
            fo_tvcsad(I0, I1, u1, u2, xi, xi + n, xi + 2 * n, xi + 3 * n, 0.85f, 0.3f, 0.125f, p->tol, w, h, warps,
                      400, log);
        free(xi);
        return 0;
    }
    if (method == 2 || method == 3) {
        if (!lab) return -1;
        fo_nltv(I0, I1, lab, u1, u2, 2.0f, 0.3f, 0.1f, w, h, warps, 400, log);
        return 0;
    }
    if (method == 6 || method == 7) {
        if (!lab) return -1;
        fo_nltvcsad(I0, I1, lab, u1, u2, 0.85f, 0.3f, 0.1f, w, h, warps, 400, log);
        return 0;
    }
    if (method == 8) {
        if (!Im1 || !chi) return -1;
        fo_tvl2occ(I0, I1, Im1, u1, u2, chi, p, w, h, warps, glb_iters, log);
        return 0;
    }
    return -1;
}
