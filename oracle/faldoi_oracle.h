/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's
 * global_faldoi hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (faldoi-ipol_b200/) never links, loads or calls it.
 *
 * PARITY PINNED: every function here is checked bit-for-bit (or to the stated
 * tolerance) against the UNMODIFIED reference compiled from /root/reference
 * (oracle/_ref/libfaldoi_ref.so, see oracle/Makefile) by tests/test_oracle_vs_ref.py,
 * and against golden vectors produced by that build (tests/golden/).
 *
 * All images are planar row-major fp32, x[j*w+i]; flow is two planes u1,u2.
 */
#ifndef FALDOI_ORACLE_H
#define FALDOI_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define FO_MAX_WARPS 64

/* per-warp log: iterations run and the last error value (what -verbose prints) */
typedef struct {
    int iters[FO_MAX_WARPS];
    float err[FO_MAX_WARPS];
} fo_log;

/* scalar parameters of the occlusion model (src/energy_structures.h:60-86) */
typedef struct {
    float lambda, theta, tau, beta, alpha, tau_u, tau_eta, tau_chi, tol, mu;
} fo_params;

/* ---- numerical primitives (src/utils.cpp, src/bicubic_interpolation.c) ---- */
void fo_centered_gradient(const float *f, float *dx, float *dy, int w, int h);
void fo_forward_gradient(const float *f, float *fx, float *fy, int w, int h);
void fo_divergence(const float *a, const float *b, float *div, int w, int h);
void fo_bicubic_warp(const float *img, const float *u, const float *v, float *out, int w, int h,
                     int border_out);

/* ---- preprocessing done by main() (src/global_faldoi.cpp:2042-2068) ---- */
void fo_rgb2gray(const float *rgb, int w, int h, float *out);
void fo_normalize3(float *i0, float *i1, float *im1, int n); /* in place, main()'s argument order */
void fo_gaussian(float *img, int w, int h, float sigma);
void fo_image_to_lab(const float *rgb, int n, float *lab);
/* full main() preprocessing: planar rgb (pd=3) or gray (pd=1) 0..255 -> I0n,I1n,I_1n */
void fo_preprocess(const float *i0, const float *i1, const float *im1, int pd, int w, int h, float *i0n,
                   float *i1n, float *im1n);
void fo_default_params(fo_params *p);

/* ---- solvers.  u1,u2 (and xi / chi) are updated in place ---- */
/* test tooling: write "warp iteration err" lines of every fo_tvl2 iteration to `path` (NULL: stop) */
void fo_set_err_trace(const char *path);
void fo_tvl2(const float *I0, const float *I1, float *u1, float *u2, float *xi11, float *xi12, float *xi21,
             float *xi22, float lambda, float theta, float tau, float tol, int w, int h, int warps,
             int max_iter, fo_log *log);
void fo_tvcsad(const float *I0, const float *I1, float *u1, float *u2, float *xi11, float *xi12, float *xi21,
               float *xi22, float lambda, float theta, float tau, float tol, int w, int h, int warps,
               int max_iter, fo_log *log);
void fo_nltv(const float *I0, const float *I1, const float *lab, float *u1, float *u2, float lambda,
             float theta, float tau, int w, int h, int warps, int max_iter, fo_log *log);
void fo_nltvcsad(const float *I0, const float *I1, const float *lab, float *u1, float *u2, float lambda,
                 float theta, float tau, int w, int h, int warps, int max_iter, fo_log *log);
void fo_tvl2occ(const float *I0, const float *I1, const float *Im1, float *u1, float *u2, float *chi,
                const fo_params *p, int w, int h, int warps, int max_iter, fo_log *log);

/* method-id dispatch with main()'s per-method constants (src/global_faldoi.cpp:2132-2167).
 * lab may be NULL for TV models; im1 may be NULL unless method==8; chi may be NULL unless method==8. */
int fo_global_solve(int method, const float *I0, const float *I1, const float *Im1, const float *lab,
                    float *u /*2*w*h*/, float *chi, const fo_params *p, int w, int h, int warps, int glb_iters,
                    fo_log *log);

#ifdef __cplusplus
}
#endif
#endif
