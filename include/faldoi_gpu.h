/* faldoi_gpu.h -- C ABI of libfaldoi_gpu.so, the B200 (sm_100a) implementation of
 * FALDOI's global variational minimisation (the `global_faldoi` pass).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * Each entry point names the reference interface it replaces (paths relative to
 * the reference repository).  All images are planar row-major fp32 (x[j*w+i]),
 * already gray / normalised / smoothed exactly as the reference's main() does
 * (src/global_faldoi.cpp:2049-2068) unless the function says otherwise.
 *
 * There is NO CPU fallback: every call returns FALDOI_ERR_CUDA (and
 * faldoi_last_error() says why) when no sm_100 device / driver is usable.
 */
#ifndef FALDOI_GPU_H
#define FALDOI_GPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* energy-model ids: src/parameters.h:5-13 */
enum {
    FALDOI_M_TVL1 = 0,
    FALDOI_M_TVL1_W = 1,
    FALDOI_M_NLTVL1 = 2,
    FALDOI_M_NLTVL1_W = 3,
    FALDOI_M_TVCSAD = 4,
    FALDOI_M_TVCSAD_W = 5,
    FALDOI_M_NLTVCSAD = 6,
    FALDOI_M_NLTVCSAD_W = 7,
    FALDOI_M_TVL1_OCC = 8
};

enum {
    FALDOI_OK = 0,
    FALDOI_ERR_ARG = 1,  /* bad argument (null pointer, size, unknown method) */
    FALDOI_ERR_CUDA = 2, /* CUDA runtime / no usable device */
    FALDOI_ERR_MEM = 3   /* allocation failed */
};

#define FALDOI_MAX_WARPS 64

/* Scalars of `Parameters` that the global step reads (src/energy_structures.h:60-86). */
typedef struct faldoi_params {
    int method;    /* val_method */
    int warps;     /* -w, default PAR_DEFAULT_NWARPS_GLOBAL = 5 (src/parameters.h:51) */
    int max_iters; /* 400 (MAX_ITERATIONS_GLOBAL) for methods 0-7; -glb_iters for method 8 */
    float lambda, theta, tau, beta, alpha, tau_u, tau_eta, tau_chi, mu, tol;
} faldoi_params;

/* What `-verbose 1` prints per warp ("Warping: k,Iter: n Error: e").
 * Iteration counts equal the reference's.  For TVL2 and TVL2-OCC the exit test is an exact maximum, so the
 * counts are bit-determined.  For TV-CSAD the test is mean |du|^2 > tol^2 (src/global_faldoi.cpp:1543): the sum is
 * taken over all pixels in double precision in a fixed order (deterministic run to run), whereas the reference
 * adds in fp32 in pixel order -- under OpenMP with a data race -- so an error within a few
 * ulp of tol^2 could leave the loop one iteration apart from a single-threaded reference run; the printed errors
 * of the CSAD / NLTV models agree to ~1e-6 relative, not bit for bit. */
typedef struct faldoi_log {
    int iters[FALDOI_MAX_WARPS];
    float err[FALDOI_MAX_WARPS];
} faldoi_log;

typedef struct faldoi_solver faldoi_solver; /* opaque; bound to one device, one frame size, one method family */

/* Defaults of init_params("", GLOBAL_STEP) (src/utils_preprocess.cpp:37-63, src/parameters.h:16-55)
 * followed by main()'s per-method overrides of lambda/theta/tau for methods 2-7
 * (src/global_faldoi.cpp:2138-2156).  glb_iters only matters for method 8. */
int faldoi_default_params(int method, int glb_iters, faldoi_params *out);

/* Same, after reading the reference's 9-line `-p` parameter file
 * (src/utils_preprocess.cpp:65-155: value <= 0 -> default, tau's > 0.25 -> default). */
int faldoi_params_from_file(const char *path, int method, int glb_iters, faldoi_params *out);

const char *faldoi_last_error(void);
int faldoi_device_count(void);

/* ---- batched solver handle -------------------------------------------------
 * One handle owns the HBM state for `batch` independent frame pairs of size
 * w x h on CUDA device `device`, plus its own stream.  Re-entrant per handle,
 * so one host thread per GPU can drive 8 GPUs.  Replaces the work-buffer
 * allocation inside each reference solver (src/global_faldoi.cpp:581-609,
 * 1195-1211, 1472-1503, 1660-1678; src/tvl2_model_occ.cpp:30-104). */
int faldoi_solver_create(faldoi_solver **out, int device, int w, int h, int method, int batch);
void faldoi_solver_destroy(faldoi_solver *s);

/* Upload pair `slot` (0 <= slot < batch).  Source pointers may be host memory (pageable or
 * pinned) or device memory of the same GPU (copies use cudaMemcpyDefault); asynchronous on the
 * handle's stream, so host buffers must stay valid until faldoi_solver_sync().
 *   I0,I1   preprocessed gray frames (w*h)
 *   Im1     preprocessed previous frame (w*h), method 8 only, else may be NULL
 *   lab     Lab image of I0 from image_to_lab (3*w*h planar), NLTV methods only, else NULL
 *   u       initial flow, 2*w*h planar (u1 | u2), e.g. the local_faldoi output
 *   chi     initial occlusion mask (w*h), method 8 only, else NULL
 * Dual variables are reset to 0 as main() does (src/global_faldoi.cpp:2116-2121). */
int faldoi_solver_upload(faldoi_solver *s, int slot, const float *I0, const float *I1, const float *Im1,
                         const float *lab, const float *u, const float *chi);

/* TV family (methods 0,1,4,5): the caller's dual variables xi11..xi22 (w*h each) as the initial duals of `slot`
 * (call after faldoi_solver_upload, which zeroes them), and the final duals after a run -- what the reference's
 * tvl2OF / tvcsad_PD read and update in place through their xi arguments (src/global_faldoi.cpp:556-573). */
int faldoi_solver_upload_xi(faldoi_solver *s, int slot, const float *xi11, const float *xi12, const float *xi21,
                            const float *xi22);
int faldoi_solver_download_xi(faldoi_solver *s, int slot, float *xi11, float *xi12, float *xi21, float *xi22);

/* Same, but from the RAW frames as iio_read_image_float_split returns them (planar, pd
 * channels, dense w*h per plane, 0..255): main()'s preprocessing -- rgb2gray, the joint
 * normalisation, gaussian(0.9) and, for the NLTV methods, image_to_lab
 * (src/global_faldoi.cpp:2042-2068) -- runs on the device.  im1 is the previous frame, or the
 * I1 pointer again when there is none (what main() does with a 2-line ims.txt).  Gray /
 * normalise / smooth give the same bits as the host code; Lab is tolerance-level. */
int faldoi_solver_upload_raw(faldoi_solver *s, int slot, const float *i0, const float *i1, const float *im1, int pd,
                             const float *u, const float *chi);
/* Inspection hook: the preprocessed frames of a slot (dense w*h planes; lab 3*w*h; any may be NULL). */
int faldoi_solver_download_frames(faldoi_solver *s, int slot, float *I0n, float *I1n, float *Im1n, float *lab);

/* Run the primal-dual minimisation on slots [0, npairs) with everything resident in HBM.
 * Replaces the bodies of tvl2OF / nltvl1_PD / tvcsad_PD / nltvcsad_PD /
 * guided_tvl2coupled_occ (dispatch: src/global_faldoi.cpp:2132-2167).
 * Asynchronous on the handle's stream; faldoi_solver_sync() waits. */
int faldoi_solver_run(faldoi_solver *s, const faldoi_params *p, int npairs);
int faldoi_solver_sync(faldoi_solver *s);

/* Download the result of pair `slot`: flow (2*w*h), and for method 8 the
 * occlusion mask chi in {0,1} (w*h); log may be NULL. */
int faldoi_solver_download(faldoi_solver *s, int slot, float *u, float *chi, faldoi_log *log);

/* NLTV models only (methods 2,3,6,7).  The default arithmetic is the reference's own -- IEEE divisions, ascending
 * slot order -- and gives bit-identical flows.  fast != 0 opts this handle into approximate divisions
 * (MUFU.RCP + FMUL, 2 ulp) and a paired slot order: about 1.7x the throughput, mean |du| ~ 1e-6 px, but the CSAD
 * variants are chaotic enough that a few pixels of a full-size frame can exceed the 1e-2 px maximum. */
int faldoi_solver_set_nltv_fast(faldoi_solver *s, int fast);

/* Milliseconds the last faldoi_solver_run spent on the device (CUDA events on the handle's stream). */
float faldoi_solver_last_run_ms(faldoi_solver *s);
/* Of that, the milliseconds between the first and last per-iteration launch of every warp
 * (i.e. excluding the bicubic warps / per-warp constants); valid after faldoi_solver_sync. */
float faldoi_solver_last_iter_ms(faldoi_solver *s);
/* Number of kernel launches issued by the last faldoi_solver_run. */
long long faldoi_solver_last_launches(faldoi_solver *s);
/* The handle's cudaStream_t (as void*) so callers can order their own work / events. */
void *faldoi_solver_stream(faldoi_solver *s);
/* Device pointer to the flow of pair `slot` (2 planes of h rows with the given pitch in floats). */
float *faldoi_solver_device_flow(faldoi_solver *s, int slot, int *pitch_floats);

/* Page-locked host memory for staging buffers (copies from / to it are asynchronous and run at full PCIe rate);
 * usable with every entry point that takes host pointers.  NULL on failure. */
void *faldoi_pinned_alloc(size_t bytes);
void faldoi_pinned_free(void *p);

/* ---- one-call host entry: H2D, solve, D2H ----------------------------------
 * The single call the host `global_faldoi` makes after preprocessing
 * (replaces src/global_faldoi.cpp:2132-2167).  u (and chi) are updated in place.
 * Re-entrant: the single-pair handle behind it is cached per host thread (and device, size, method), so one
 * host thread per GPU can drive 8 GPUs concurrently. */
int faldoi_global_solve(int device, const faldoi_params *p, int w, int h, const float *I0, const float *I1,
                        const float *Im1, const float *lab, float *u, float *chi, faldoi_log *log);

/* The same from raw frames (device-side preprocessing, see faldoi_solver_upload_raw): everything
 * main() does between reading the files and saving the flow. */
int faldoi_global_solve_raw(int device, const faldoi_params *p, int w, int h, int pd, const float *i0, const float *i1,
                            const float *im1, float *u, float *chi, faldoi_log *log);

/* ---- per-solver mirrors of the reference's in-process signatures -----------
 * Same argument order and in-place semantics as the reference functions (u1,u2 AND the xi arrays of tvl2OF /
 * tvcsad_PD are read as the initial state and hold the final state on return; xi pointers may be NULL = start
 * from zero); the only addition is the int status.  `verbose` prints the reference's per-warp
 * line to stderr (methods 0,4) / stdout (methods 2,6).
 *   tvl2OF       src/global_faldoi.cpp:556-573
 *   tvcsad_PD    src/global_faldoi.cpp:1449-1466
 *   nltvl1_PD    src/global_faldoi.cpp:1177-1191
 *   nltvcsad_PD  src/global_faldoi.cpp:1642-1656 */
int faldoi_tvl2OF(const float *I0, float *I1, float *u1, float *u2, float *xi11, float *xi12, float *xi21,
                  float *xi22, float lambda, float theta, float tau, float tol_OF, int nx, int ny, int warps,
                  int verbose);
int faldoi_tvcsad_PD(const float *I0, float *I1, float *xi11, float *xi12, float *xi21, float *xi22,
                     float lambda, float theta, float tau, float tol_OF, int nx, int ny, int warps, int verbose,
                     float *u1, float *u2);
int faldoi_nltvl1_PD(const float *I0, float *I1, float *a, int pd, float lambda, float theta, float tau, int w,
                     int h, int warps, int verbose, float *u1, float *u2);
int faldoi_nltvcsad_PD(const float *I0, float *I1, float *a, int pd, float lambda, float theta, float tau, int w,
                       int h, int warps, int verbose, float *u1, float *u2);
/* guided_tvl2coupled_occ (src/tvl2_model_occ.cpp:492-502) with the patch set to
 * the whole image and step_algorithm = GLOBAL_STEP, scalars taken from `p`
 * instead of ofD->params.  eta and div_u start at 0 (the reference leaves them
 * uninitialised). */
int faldoi_guided_tvl2coupled_occ(const float *I0, const float *I1, const float *I_1, float *u1, float *u2,
                                  float *chi, const faldoi_params *p, int nx, int ny, int verbose);

/* ---- row stripes: ONE large frame pair over several GPUs (>= 4K frames) ------------
 * The frame is cut into `nstripes` contiguous row blocks, stripe k on CUDA device
 * devices[k] (devices may repeat, e.g. to exercise the path on one GPU).  Every launch of the
 * iteration kernel stores the rows next to a stripe boundary directly into the neighbour GPU's
 * halo rows over NVLink peer mappings and signals it through a flag word in its memory; the exit
 * test uses the maximum over the whole frame (gathered once per block of launches, with a
 * checkpoint to roll back to), so results (flow, per-warp iteration counts, errors) are identical
 * to the single-GPU solve.  Needs peer access between the devices and stream memory operations.
 * Implemented for TVL2 (methods 0,1).  The reference has no counterpart: its tvl2OF
 * (src/global_faldoi.cpp:556) is single-node OpenMP. */
typedef struct faldoi_stripes faldoi_stripes;
int faldoi_stripe_rows(int h, int nstripes, int k, int *row0, int *row1); /* rows [row0,row1) owned by stripe k */
int faldoi_stripes_create(faldoi_stripes **out, int nstripes, const int *devices, int w, int h, int method);
void faldoi_stripes_destroy(faldoi_stripes *g);
/* full-frame arrays (host, or device memory reachable from every stripe's GPU): I0, I1 (w*h), u (2*w*h) */
int faldoi_stripes_upload(faldoi_stripes *g, const float *I0, const float *I1, const float *u);
int faldoi_stripes_run(faldoi_stripes *g, const faldoi_params *p); /* synchronous: one host thread per stripe */
int faldoi_stripes_download(faldoi_stripes *g, float *u, faldoi_log *log);
float faldoi_stripes_last_run_ms(faldoi_stripes *g);
long long faldoi_stripes_last_launches(faldoi_stripes *g);

/* ---- device-side preprocessing helpers (row "next" of the scope table) ---- */
/* centered_gradient (src/utils.cpp:367-423) and bicubic_interpolation_warp
 * (src/bicubic_interpolation.c:245-266) on host arrays, computed on the GPU. */
int faldoi_centered_gradient(int device, const float *in, float *dx, float *dy, int nx, int ny);
int faldoi_bicubic_warp(int device, const float *in, const float *u, const float *v, float *out, int nx, int ny,
                        int border_out);

/* Numerical self-test: the shared-reciprocal division used by the TVL2 dual projection against
 * IEEE division on n pseudo-random / structured operand pairs; *mismatches must come back 0. */
int faldoi_selftest_division(int device, unsigned long long n, unsigned long long seed, unsigned long long *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* FALDOI_GPU_H */
