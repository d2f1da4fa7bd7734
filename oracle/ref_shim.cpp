// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// extern "C" handles onto the UNMODIFIED reference solvers, so tests can call
// them in-process (ctypes) on arbitrary arrays.  This file is compiled by
// oracle/Makefile together with the reference sources where they lie under
// /root/reference/src; the result goes to oracle/_ref/libfaldoi_ref.so
// (git-ignored).  No reference source is copied: we only declare the
// prototypes of the functions we bind, citing where each one is defined.
//
//   tvl2OF                  src/global_faldoi.cpp:556
//   nltvl1_PD               src/global_faldoi.cpp:1177
//   tvcsad_PD               src/global_faldoi.cpp:1449
//   nltvcsad_PD             src/global_faldoi.cpp:1642
//   image_to_lab            src/global_faldoi.cpp:906
//   rgb2gray                src/global_faldoi.cpp:1820
//   main (renamed by -Dmain=ref_global_faldoi_main)  src/global_faldoi.cpp:1846
//   guided_tvl2coupled_occ  src/tvl2_model_occ.cpp:492
//   initialize_auxiliar_stuff / free_auxiliar_stuff  src/energy_model.cpp:202,240
//   init_Optical_Flow_Data  src/energy_model.cpp:161
//   init_params             src/utils_preprocess.cpp:37
//   centered_gradient, forward_gradient, divergence, gaussian,
//   image_normalization_3   src/utils.cpp:367,285,239,521,743
//   bicubic_interpolation_warp  src/bicubic_interpolation.c:245
#include <cstring>
#include <string>

#include "energy_structures.h"
#include "energy_model.h"
#include "tvl2_model_occ.h"
#include "utils.h"
#include "utils_preprocess.h"
#include "parameters.h"
extern "C" {
#include "bicubic_interpolation.h"
}

void tvl2OF(const float *I0, float *I1, float *u1, float *u2, float *xi11, float *xi12, float *xi21,
            float *xi22, const float lambda, const float theta, const float tau, const float tol_OF,
            const int nx, const int ny, const int warps, const bool verbose);
void nltvl1_PD(const float *I0, float *I1, float *a, int pd, const float lambda, const float theta,
               const float tau, const int w, const int h, const int warps, const bool verbose, float *u1,
               float *u2);
void tvcsad_PD(const float *I0, float *I1, float *xi11, float *xi12, float *xi21, float *xi22,
               const float lambda, const float theta, const float tau, const float tol_OF, const int nx,
               const int ny, const int warps, const bool verbose, float *u1, float *u2);
void nltvcsad_PD(const float *I0, float *I1, float *a, int pd, const float lambda, const float theta,
                 const float tau, const int w, const int h, const int warps, const bool verbose, float *u1,
                 float *u2);
void image_to_lab(float *in, int size, float *out);
void rgb2gray(float *in, int w, int h, float *out);
int ref_global_faldoi_main(int argc, char *argv[]);

extern "C" {

void ref_tvl2OF(const float *I0, float *I1, float *u1, float *u2, float *xi11, float *xi12, float *xi21,
                float *xi22, float lambda, float theta, float tau, float tol, int nx, int ny, int warps,
                int verbose) {
    tvl2OF(I0, I1, u1, u2, xi11, xi12, xi21, xi22, lambda, theta, tau, tol, nx, ny, warps, verbose != 0);
}

void ref_nltvl1_PD(const float *I0, float *I1, float *lab, int pd, float lambda, float theta, float tau,
                   int w, int h, int warps, int verbose, float *u1, float *u2) {
    nltvl1_PD(I0, I1, lab, pd, lambda, theta, tau, w, h, warps, verbose != 0, u1, u2);
}

void ref_tvcsad_PD(const float *I0, float *I1, float *xi11, float *xi12, float *xi21, float *xi22,
                   float lambda, float theta, float tau, float tol, int nx, int ny, int warps, int verbose,
                   float *u1, float *u2) {
    tvcsad_PD(I0, I1, xi11, xi12, xi21, xi22, lambda, theta, tau, tol, nx, ny, warps, verbose != 0, u1, u2);
}

void ref_nltvcsad_PD(const float *I0, float *I1, float *lab, int pd, float lambda, float theta, float tau,
                     int w, int h, int warps, int verbose, float *u1, float *u2) {
    nltvcsad_PD(I0, I1, lab, pd, lambda, theta, tau, w, h, warps, verbose != 0, u1, u2);
}

// Method 8 exactly as main() drives it (src/global_faldoi.cpp:2022-2032,
// 2093-2122, 2165): default/`-p` parameters, whole-image patch, GLOBAL_STEP.
// u is 2*w*h (u1|u2) in/out, chi is w*h in/out.  The reference never
// initialises eta1/eta2/div_u on this path; the defined
// target is eta = div_u = 0, so this shim zeroes those three work buffers
// after the reference allocated them.
void ref_tvl2occ(const float *I0, const float *I1, const float *I_1, float *u, float *chi,
                 const char *params_file, int w, int h, int warps, int iters, int verbose) {
    Parameters params = init_params(std::string(params_file ? params_file : ""), GLOBAL_STEP);
    params.w = w;
    params.h = h;
    params.warps = warps;
    params.val_method = M_TVL1_OCC;
    params.iterations_of = iters;
    params.verbose = verbose != 0;
    OpticalFlowData ofD = init_Optical_Flow_Data(params);
    const size_t n = (size_t)w * h;
    std::memcpy(ofD.u1, u, 2 * n * sizeof(float));
    std::memcpy(ofD.chi, chi, n * sizeof(float));
    SpecificOFStuff stuff{};
    initialize_auxiliar_stuff(stuff, ofD, w, h);
    std::memset(stuff.tvl2_occ.eta1, 0, n * sizeof(float));
    std::memset(stuff.tvl2_occ.eta2, 0, n * sizeof(float));
    std::memset(stuff.tvl2_occ.div_u, 0, n * sizeof(float));
    for (size_t i = 0; i < n; i++)
        stuff.tvl2_occ.xi11[i] = stuff.tvl2_occ.xi12[i] = stuff.tvl2_occ.xi21[i] = stuff.tvl2_occ.xi22[i] = 0.f;
    PatchIndexes index{};
    index.ii = 0;
    index.ij = 0;
    index.ei = w;
    index.ej = h;
    float ener = 0.f;
    guided_tvl2coupled_occ(I0, I1, I_1, &ofD, &stuff.tvl2_occ, &ener, index, w, h);
    std::memcpy(u, ofD.u1, 2 * n * sizeof(float));
    std::memcpy(chi, ofD.chi, n * sizeof(float));
    free_auxiliar_stuff(&stuff, &ofD);
    delete[] ofD.u1;
    delete[] ofD.u1_ba;
    delete[] ofD.chi;
}

void ref_image_to_lab(float *rgb, int size, float *out) { image_to_lab(rgb, size, out); }
void ref_rgb2gray(float *rgb, int w, int h, float *out) { rgb2gray(rgb, w, h, out); }
void ref_image_normalization_3(const float *a, const float *b, const float *c, float *an, float *bn,
                               float *cn, int size) {
    image_normalization_3(a, b, c, an, bn, cn, size);
}
void ref_gaussian(float *I, int w, int h, float sigma) { gaussian(I, w, h, sigma); }
void ref_centered_gradient(const float *in, float *dx, float *dy, int nx, int ny) {
    centered_gradient(in, dx, dy, nx, ny);
}
void ref_forward_gradient(const float *f, float *fx, float *fy, int nx, int ny) {
    forward_gradient(f, fx, fy, nx, ny);
}
void ref_divergence(const float *v1, const float *v2, float *div, int nx, int ny) {
    divergence(v1, v2, div, nx, ny);
}
void ref_bicubic_warp(const float *in, const float *u, const float *v, float *out, int nx, int ny,
                      int border_out) {
    bicubic_interpolation_warp(in, u, v, out, nx, ny, border_out != 0);
}
// The 9 scalars init_params() yields for a given `-p` file ("" = defaults):
// lambda, theta, tau, beta, alpha, tau_u, tau_eta, tau_chi, mu, then tol_OF.
void ref_init_params(const char *params_file, float *out10) {
    Parameters p = init_params(std::string(params_file ? params_file : ""), GLOBAL_STEP);
    out10[0] = p.lambda;
    out10[1] = p.theta;
    out10[2] = p.tau;
    out10[3] = p.beta;
    out10[4] = p.alpha;
    out10[5] = p.tau_u;
    out10[6] = p.tau_eta;
    out10[7] = p.tau_chi;
    out10[8] = p.mu;
    out10[9] = p.tol_OF;
}
int ref_global_faldoi(int argc, char **argv) { return ref_global_faldoi_main(argc, argv); }

}  // extern "C"
