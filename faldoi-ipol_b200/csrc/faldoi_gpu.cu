// libfaldoi_gpu.so -- C ABI (include/faldoi_gpu.h) over the sm_100a kernels.
// Build: see faldoi-ipol_b200/build.py  (nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false ...)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>

#include "solver.h"
#include "tv_kernels.cuh"
#include "tma.cuh"
#include "warp_kernels.cuh"
#include "nltv_kernels.cuh"
#include "occ_kernels.cuh"
#include "preprocess_kernels.cuh"

using namespace faldoi;

static int occ_upload(faldoi_solver *s, int slot, const float *Im1, const float *u, const float *chi);
static int run_nltv(faldoi_solver *s, const faldoi_params *p, int npairs);
static int run_occ(faldoi_solver *s, const faldoi_params *p, int npairs);
static int up2d(faldoi_solver *s, float *dst_plane, const float *src);
static dim3 grid2d(const Geo &g, dim3 block, int npairs);

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
namespace faldoi {
static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
bool cuda_ok(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}
}  // namespace faldoi

extern "C" const char *faldoi_last_error(void) { return g_err.c_str(); }

extern "C" int faldoi_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// ---------------------------------------------------------------------------
// parameters: init_params (src/utils_preprocess.cpp:37-157) + main()'s overrides
// ---------------------------------------------------------------------------
static void base_defaults(faldoi_params *p) {
    // src/parameters.h:20-31 (double literals narrowed to float, as the reference's `params.x = PAR_...`)
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
    p->warps = 5;
}

static int finish_params(int method, int glb_iters, faldoi_params *p) {
    if (method < 0 || method > 8) {
        set_error("unknown method id");
        return FALDOI_ERR_ARG;
    }
    p->method = method;
    // src/global_faldoi.cpp:2138-2156: methods 2-7 overwrite lambda/theta/tau after init_params
    if (method == FALDOI_M_NLTVCSAD || method == FALDOI_M_NLTVCSAD_W) {
        p->lambda = 0.85;
        p->theta = 0.3;
        p->tau = 0.1;
    } else if (method == FALDOI_M_NLTVL1 || method == FALDOI_M_NLTVL1_W) {
        p->lambda = 2.0;
        p->theta = 0.3;
        p->tau = 0.1;
    } else if (method == FALDOI_M_TVCSAD || method == FALDOI_M_TVCSAD_W) {
        p->lambda = 0.85;
        p->theta = 0.3;
        p->tau = 0.125;
    }
    // methods 0-7 loop to the compile-time MAX_ITERATIONS_GLOBAL; only method 8 reads -glb_iters
    p->max_iters = (method == FALDOI_M_TVL1_OCC) ? glb_iters : 400;
    return FALDOI_OK;
}

extern "C" int faldoi_default_params(int method, int glb_iters, faldoi_params *out) {
    if (!out) return FALDOI_ERR_ARG;
    base_defaults(out);
    return finish_params(method, glb_iters, out);
}

extern "C" int faldoi_params_from_file(const char *path, int method, int glb_iters, faldoi_params *out) {
    if (!out) return FALDOI_ERR_ARG;
    base_defaults(out);
    if (path && path[0]) {
        std::ifstream in(path);
        std::string line;
        float v[9];
        for (int i = 0; i < 9; i++) {
            // the reference calls std::stof on each line and dies on a missing / short file
            if (!std::getline(in, line)) {
                set_error(std::string("parameter file missing or shorter than 9 lines: ") + path);
                return FALDOI_ERR_ARG;
            }
            try {
                v[i] = std::stof(line);
            } catch (...) {
                set_error(std::string("parameter file: not a number: ") + line);
                return FALDOI_ERR_ARG;
            }
        }
        const faldoi_params d = *out;
        out->lambda = v[0] <= 0 ? d.lambda : v[0];
        out->theta = v[1] <= 0 ? d.theta : v[1];
        out->tau = (v[2] <= 0 || v[2] > 0.25) ? d.tau : v[2];
        out->beta = v[3] <= 0 ? d.beta : v[3];
        out->alpha = v[4] <= 0 ? d.alpha : v[4];
        out->tau_u = (v[5] <= 0 || v[5] > 0.25) ? d.tau_u : v[5];
        out->tau_eta = (v[6] <= 0 || v[6] > 0.25) ? d.tau_eta : v[6];
        out->tau_chi = (v[7] <= 0 || v[7] > 0.25) ? d.tau_chi : v[7];
        out->mu = v[8] <= 0 ? d.mu : v[8];
    }
    return finish_params(method, glb_iters, out);
}

// ---------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------
// Every device array carries a 1 KiB guard band on both sides: the tile kernels copy whole
// 16-byte-aligned row segments including a 4-pixel halo left of column 0 / right of the last
// tile, which for the first / last row of an array falls just outside it (never used in
// arithmetic, but it must be addressable).
float *faldoi_solver::dmalloc(size_t nfloats) {
    const size_t guard = 256;  // floats
    void *p = nullptr;
    if (!cuda_ok(cudaMalloc(&p, (nfloats + 2 * guard) * sizeof(float)), "cudaMalloc")) return nullptr;
    if (!cuda_ok(cudaMemsetAsync(p, 0, (nfloats + 2 * guard) * sizeof(float), stream), "cudaMemset")) return nullptr;
    allocs.push_back(p);
    return (float *)p + guard;
}

int faldoi_solver::phase_mark() {
    if (phase_used == (int)phase_ev.size()) {
        cudaEvent_t e;
        FALDOI_CUDA(cudaEventCreate(&e));
        phase_ev.push_back(e);
    }
    FALDOI_CUDA(cudaEventRecord(phase_ev[phase_used++], stream));
    return FALDOI_OK;
}

int faldoi_solver::alloc_err(int max_iters) {
    if (max_iters <= err_cap) return FALDOI_OK;
    // (old arrays stay in `allocs` until destroy; growth happens at most a few times)
    err_max = (unsigned *)dmalloc((size_t)B * max_iters);
    err_sum = (double *)dmalloc((size_t)B * max_iters * 2);
    if (!err_max || !err_sum) return FALDOI_ERR_MEM;
    err_cap = max_iters;
    return FALDOI_OK;
}

// ---------------------------------------------------------------------------
// TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no -lcuda)
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_plane_map(CUtensorMap *out, float *base, const Geo &g, size_t nplanes, int box_w, int box_h) {
    static EncodeTiledFn encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return (EncodeTiledFn)fn;
    }();
    if (!encode) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return FALDOI_ERR_CUDA;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)g.pitch, (cuuint64_t)g.h, (cuuint64_t)nplanes};
    const cuuint64_t strides[2] = {(cuuint64_t)g.pitch * sizeof(float), (cuuint64_t)g.plane * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return FALDOI_ERR_CUDA;
    }
    return FALDOI_OK;
}

static int make_tile_maps(faldoi_solver *s) {
    const Geo &g = s->g;
    const size_t nstate = (size_t)2 * ST_COUNT * g.B;
    float *c0 = s->rho_c;
    int rc;
    if (method_is_csad(s->method)) {  // boxes of the two-iteration TV-CSAD kernel
        if ((rc = make_plane_map(&s->cmaps.ub, s->state, g, nstate, C2_PW, C2_UB_ROWS))) return rc;
        if ((rc = make_plane_map(&s->cmaps.xi, s->state, g, nstate, C2_PW, C2_XI_ROWS))) return rc;
        if ((rc = make_plane_map(&s->cmaps.pl, s->state, g, nstate, C2_PW, C2_PL_ROWS))) return rc;
        if ((rc = make_plane_map(&s->cmaps.sc, s->scale, g, g.B, C2_PW, C2_PL_ROWS))) return rc;
        if ((rc = make_plane_map(&s->cmaps.ix, s->Ix, g, g.B, C2_PW, C2_PL_ROWS))) return rc;
        if ((rc = make_plane_map(&s->cmaps.iy, s->Iy, g, g.B, C2_PW, C2_PL_ROWS))) return rc;
        if ((rc = make_plane_map(&s->cmaps.t1, s->csad_t1, g, g.B, C2_PW, C2_PL_ROWS))) return rc;
        if ((rc = make_plane_map(&s->cmaps.t2, s->csad_t2, g, g.B, C2_PW, C2_PL_ROWS))) return rc;
        if ((rc = make_plane_map(&s->cmaps.i0, s->I0, g, g.B, C2_TP, C2_T_ROWS))) return rc;
        if ((rc = make_plane_map(&s->cmaps.iw, s->I1w, g, g.B, C2_TP, C2_T_ROWS))) return rc;
    }
    if (!method_is_csad(s->method)) {  // boxes of the two-iteration kernel (2-pixel apron)
        if ((rc = make_plane_map(&s->maps2.ub, s->state, g, nstate, T2_PW, T2_UB_ROWS))) return rc;
        if ((rc = make_plane_map(&s->maps2.xi, s->state, g, nstate, T2_PW, T2_XI_ROWS))) return rc;
        if ((rc = make_plane_map(&s->maps2.pl, s->state, g, nstate, T2_PW, T2_PL_ROWS))) return rc;
        if ((rc = make_plane_map(&s->maps2.c0, c0, g, g.B, T2_PW, T2_PL_ROWS))) return rc;
        if ((rc = make_plane_map(&s->maps2.ix, s->Ix, g, g.B, T2_PW, T2_PL_ROWS))) return rc;
        if ((rc = make_plane_map(&s->maps2.iy, s->Iy, g, g.B, T2_PW, T2_PL_ROWS))) return rc;
    }
    return FALDOI_OK;
}

struct StripeGeo {
    int y_off, hg, own_lo, own_hi;
};
static int create_internal(faldoi_solver **out, int device, int w, int h, int method, int batch, const StripeGeo *sg);

extern "C" int faldoi_solver_create(faldoi_solver **out, int device, int w, int h, int method, int batch) {
    return create_internal(out, device, w, h, method, batch, nullptr);
}

static int create_internal(faldoi_solver **out, int device, int w, int h, int method, int batch, const StripeGeo *sg) {
    if (!out || w < 2 || h < 2 || batch < 1 || method < 0 || method > 8) {
        set_error("faldoi_solver_create: bad argument");
        return FALDOI_ERR_ARG;
    }
    int ndev = 0;
    FALDOI_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) {
        set_error("faldoi_solver_create: no such CUDA device");
        return FALDOI_ERR_CUDA;
    }
    FALDOI_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    FALDOI_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(std::string("libfaldoi_gpu is built for sm_100a only; device is ") + prop.name);
        return FALDOI_ERR_CUDA;
    }
    faldoi_solver *s = new faldoi_solver();
    s->device = device;
    s->method = method;
    s->B = batch;
    s->g = make_geo(w, h, (w + 31) / 32 * 32, batch);
    s->nl_parity.assign(batch, 0);
    s->i1_plane = s->g.plane;
    if (sg) {  // row stripe of a taller frame: I1 and its gradients stay full frames (the warp samples any row)
        s->g.y_off = sg->y_off, s->g.hg = sg->hg, s->g.own_lo = sg->own_lo, s->g.own_hi = sg->own_hi;
        s->i1_plane = (size_t)s->g.pitch * sg->hg;
    }
    const size_t P = s->g.plane, B = batch;
    auto fail = [&](int code) {
        faldoi_solver_destroy(s);
        return code;
    };
    if (!cuda_ok(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking), "cudaStreamCreate")) return fail(FALDOI_ERR_CUDA);
    if (!cuda_ok(cudaEventCreate(&s->ev0), "cudaEventCreate") || !cuda_ok(cudaEventCreate(&s->ev1), "cudaEventCreate"))
        return fail(FALDOI_ERR_CUDA);

#define ALLOC(ptr, n)                                  \
    do {                                               \
        (ptr) = s->dmalloc((size_t)(n));               \
        if (!(ptr)) return fail(FALDOI_ERR_MEM);       \
    } while (0)

    ALLOC(s->I0, B * P);
    ALLOC(s->I1, B * s->i1_plane);
    ALLOC(s->I1x, B * s->i1_plane);
    ALLOC(s->I1y, B * s->i1_plane);
    ALLOC(s->packed, B * 3 * (size_t)w * h);
    s->parity = (int *)s->dmalloc(B);
    s->log_iters = (int *)s->dmalloc(B * FALDOI_MAX_WARPS);
    s->log_err = s->dmalloc(B * FALDOI_MAX_WARPS);
    if (!s->parity || !s->log_iters || !s->log_err) return fail(FALDOI_ERR_MEM);
    if (s->alloc_err(400) != FALDOI_OK) return fail(FALDOI_ERR_MEM);
    s->d_active = (int *)s->dmalloc(4);
    if (!s->d_active || !cuda_ok(cudaHostAlloc((void **)&s->h_active, 4 * sizeof(int), cudaHostAllocDefault), "cudaHostAlloc"))
        return fail(FALDOI_ERR_MEM);
    for (int k = 0; k < 4; k++)
        if (!cuda_ok(cudaEventCreateWithFlags(&s->chunk_ev[k], cudaEventDisableTiming), "cudaEventCreate")) return fail(FALDOI_ERR_CUDA);

    const Family fam = method_family(method);
    if (fam == FAM_TV || fam == FAM_NLTV) {
        s->set_stride = (size_t)ST_COUNT * B * P;
        ALLOC(s->state, 2 * s->set_stride);
        ALLOC(s->Ix, B * P);
        ALLOC(s->Iy, B * P);
        if (method_is_csad(method)) {
            ALLOC(s->scale, B * P);
            ALLOC(s->I1w, B * P);
            if (fam == FAM_TV) {  // TV-CSAD: rank order of the neighbour residuals (tv_csad2_kernel.cuh)
                ALLOC(s->csad_t1, B * P);
                ALLOC(s->csad_t2, B * P);
                s->csad_perm = (unsigned *)s->dmalloc((size_t)C2_WORDS * B * P);
                const size_t ncta = (size_t)((s->g.pitch + C2_W - 1) / C2_W) * ((h + C2_H - 1) / C2_H);
                s->csad_partial = (double *)s->dmalloc(2 * (2 * B * ncta));
                if (!s->csad_perm || !s->csad_partial) return fail(FALDOI_ERR_MEM);
            } else {  // NLTV-CSAD: two-level sorted residual table (csad_select)
                ALLOC(s->csad_blk, CSAD_FLOATS * B * P);
                ALLOC(s->csad_sep, CSAD_SEPS * B * P);
            }
        } else {
            ALLOC(s->rho_c, B * P);
        }
    }
    if (fam == FAM_TV && make_tile_maps(s) != FALDOI_OK) return fail(FALDOI_ERR_CUDA);
    if (fam == FAM_NLTV) {
        ALLOC(s->lab, 3 * B * P);
        ALLOC(s->wgt, (size_t)NL_SLOTS * B * P);
        ALLOC(s->wt, B * P);
        ALLOC(s->rwt, B * P);
        s->dual_set_stride = (size_t)2 * NL_SLOTS * B * P;
        ALLOC(s->dual, 2 * s->dual_set_stride);
        if (make_plane_map(&s->nlmaps.dual, s->dual, s->g, (size_t)4 * NL_SLOTS * B, NLT_W, NLT_H) ||
            make_plane_map(&s->nlmaps.rec, s->dual, s->g, (size_t)4 * NL_SLOTS * B, NLT_PW, NLT_H) ||
            make_plane_map(&s->nlmaps.wgt, s->wgt, s->g, (size_t)NL_SLOTS * B, NLT_W, NLT_H) ||
            make_plane_map(&s->nlmaps.ub, s->state, s->g, (size_t)2 * ST_COUNT * B, NLT_PW, NLT_AR) ||
            make_plane_map(&s->nlmaps.rwt, s->rwt, s->g, B, NLT_PW, NLT_AR) ||
            make_plane_map(&s->nlmaps.wt, s->wt, s->g, B, NLT_PW, NLT_AR))
            return fail(FALDOI_ERR_CUDA);
    }
    if (fam == FAM_OCC) {
        ALLOC(s->occ, (size_t)OCC_PLANES * B * P);
    }
#undef ALLOC
    // function attributes are per device: opt the tile kernels into 97 KB of dynamic shared memory here
    if (!cuda_ok(cudaFuncSetAttribute(tv_csad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Csad2Smem) + 128),
                 "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(tv_tile2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile2Smem) + 128),
                 "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(tv_tile2_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Tile2Smem) + 128),
                 "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(nltv_tile_kernel<DATA_TVL1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NlTileSmem) + 128),
                 "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(nltv_tile_kernel<DATA_CSAD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NlTileSmem) + 128),
                 "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(nltv_tile_kernel<DATA_TVL1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NlTileSmem) + 128),
                 "cudaFuncSetAttribute") ||
        !cuda_ok(cudaFuncSetAttribute(nltv_tile_kernel<DATA_CSAD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NlTileSmem) + 128),
                 "cudaFuncSetAttribute"))
        return fail(FALDOI_ERR_CUDA);
    if (!cuda_ok(cudaStreamSynchronize(s->stream), "cudaStreamSynchronize")) return fail(FALDOI_ERR_CUDA);
    *out = s;
    return FALDOI_OK;
}

extern "C" void faldoi_solver_destroy(faldoi_solver *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (void *p : s->allocs) cudaFree(p);
    for (cudaEvent_t e : s->phase_ev) cudaEventDestroy(e);
    for (int k = 0; k < 4; k++)
        if (s->chunk_ev[k]) cudaEventDestroy(s->chunk_ev[k]);
    if (s->h_active) cudaFreeHost(s->h_active);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

static int up2d(faldoi_solver *s, float *dst_plane, const float *src) {
    FALDOI_CUDA(cudaMemcpy2DAsync(dst_plane, s->g.pitch * sizeof(float), src, s->g.w * sizeof(float),
                                  s->g.w * sizeof(float), s->g.h, cudaMemcpyDefault, s->stream));
    return FALDOI_OK;
}

extern "C" int faldoi_solver_upload(faldoi_solver *s, int slot, const float *I0, const float *I1, const float *Im1,
                                    const float *lab, const float *u, const float *chi) {
    if (!s || slot < 0 || slot >= s->B || !I0 || !I1 || !u) {
        set_error("faldoi_solver_upload: bad argument");
        return FALDOI_ERR_ARG;
    }
    const Family fam = method_family(s->method);
    if ((fam == FAM_NLTV && !lab) || (fam == FAM_OCC && (!Im1 || !chi))) {
        set_error("faldoi_solver_upload: this method needs lab (NLTV) / Im1 and chi (occlusions)");
        return FALDOI_ERR_ARG;
    }
    FALDOI_CUDA(cudaSetDevice(s->device));
    const size_t P = s->g.plane, B = s->B, n = (size_t)s->g.w * s->g.h;
    int rc;
    if ((rc = up2d(s, s->I0 + slot * P, I0))) return rc;
    if ((rc = up2d(s, s->I1 + slot * P, I1))) return rc;
    if (fam == FAM_TV || fam == FAM_NLTV) {
        // flow into set 0, parity 0, duals zeroed (src/global_faldoi.cpp:2116-2121; NLTV sc = 0 :1027)
        float *set0 = s->state;
        if ((rc = up2d(s, set0 + (ST_U1 * B + slot) * P, u))) return rc;
        if ((rc = up2d(s, set0 + (ST_U2 * B + slot) * P, u + n))) return rc;
        for (int k = ST_XI11; k <= ST_XI22; k++)
            FALDOI_CUDA(cudaMemsetAsync(set0 + (k * B + slot) * P, 0, P * sizeof(float), s->stream));
        FALDOI_CUDA(cudaMemsetAsync(s->parity + slot, 0, sizeof(int), s->stream));
    }
    if (fam == FAM_NLTV) {
        for (int c = 0; c < 3; c++)
            if ((rc = up2d(s, s->lab + (c * B + slot) * P, lab + c * n))) return rc;
        for (int k = 0; k < 2 * NL_SLOTS; k++)
            FALDOI_CUDA(cudaMemsetAsync(s->dual + (k * B + slot) * P, 0, P * sizeof(float), s->stream));
        s->nl_parity[slot] = 0;
    }
    if (fam == FAM_OCC) {
        if ((rc = occ_upload(s, slot, Im1, u, chi))) return rc;
    }
    return FALDOI_OK;
}

// ---------------------------------------------------------------------------
// upload with device-side preprocessing (gray, joint normalisation, Gaussian, Lab)
// ---------------------------------------------------------------------------
static GaussTaps presmoothing_taps() {
    // coefficients exactly as gaussian() builds them on the host (src/utils.cpp:541-553), sigma = 0.9
    GaussTaps t{};
    const float sigma = 0.90f, den = 2 * sigma * sigma;
    t.taps = (int)(5 * sigma) + 1;
    for (int i = 0; i < t.taps; i++) t.k[i] = (float)(1 / (sigma * sqrt(2.0 * 3.1415926)) * expf((float)(-i * i) / den));
    float norm = 0;
    for (int i = 0; i < t.taps; i++) norm += t.k[i];
    norm *= 2;
    norm -= t.k[0];
    for (int i = 0; i < t.taps; i++) t.k[i] /= norm;
    return t;
}

extern "C" int faldoi_solver_upload_raw(faldoi_solver *s, int slot, const float *i0, const float *i1, const float *im1, int pd,
                                        const float *u, const float *chi) {
    if (!s || slot < 0 || slot >= s->B || !i0 || !i1 || !im1 || !u || pd < 1) {
        set_error("faldoi_solver_upload_raw: bad argument");
        return FALDOI_ERR_ARG;
    }
    const Family fam = method_family(s->method);
    if (fam == FAM_NLTV && pd < 3) {
        set_error("the NLTV models need a colour (3-channel) first frame");
        return FALDOI_ERR_ARG;
    }
    if (fam == FAM_OCC && !chi) {
        set_error("faldoi_solver_upload_raw: method 8 needs chi");
        return FALDOI_ERR_ARG;
    }
    const Geo g = s->g;
    if (g.h < 5 || g.w < 5) {
        set_error("gaussian: sigma too large for the image");
        return FALDOI_ERR_ARG;
    }
    FALDOI_CUDA(cudaSetDevice(s->device));
    const size_t n = (size_t)g.w * g.h, P = g.plane, B = s->B, need = 3 * (size_t)pd * n;
    if (need > s->raw_cap) {
        s->raw_stage = s->dmalloc(need);
        if (!s->raw_stage) return FALDOI_ERR_MEM;
        s->raw_cap = need;
    }
    if (!s->pp_tmp) {
        s->pp_tmp = s->dmalloc(P);
        s->pp_im1 = s->dmalloc(P);
        s->pp_mm = (unsigned *)s->dmalloc(8);
        if (!s->pp_tmp || !s->pp_im1 || !s->pp_mm) return FALDOI_ERR_MEM;
    }
    const float *src[3] = {i0, i1, im1};
    for (int k = 0; k < 3; k++)
        FALDOI_CUDA(cudaMemcpyAsync(s->raw_stage + k * pd * n, src[k], pd * n * sizeof(float), cudaMemcpyDefault, s->stream));
    static const unsigned mm_init[6] = {0xffffffffu, 0u, 0xffffffffu, 0u, 0xffffffffu, 0u};
    FALDOI_CUDA(cudaMemcpyAsync(s->pp_mm, mm_init, sizeof(mm_init), cudaMemcpyHostToDevice, s->stream));
    float *d_i0 = s->I0 + slot * P, *d_i1 = s->I1 + slot * P;
    float *d_im1 = (fam == FAM_OCC) ? s->occ + (OC_IM1 * B + slot) * P : s->pp_im1;
    float *dst[3] = {d_i0, d_i1, d_im1};
    const dim3 blk(32, 8);
    Geo g1 = g;
    g1.B = 1;
    const dim3 grd = grid2d(g1, blk, 1);
    for (int k = 0; k < 3; k++) gray_minmax_kernel<<<grd, blk, 0, s->stream>>>(s->raw_stage + k * pd * n, pd, dst[k], s->pp_mm + 2 * k, g1);
    normalize3_kernel<<<grd, blk, 0, s->stream>>>(d_i0, d_i1, d_im1, s->pp_mm, g1);
    static const GaussTaps taps = presmoothing_taps();
    for (int k = 0; k < (fam == FAM_OCC ? 3 : 2); k++) {
        gaussian_pass_kernel<<<grd, blk, 0, s->stream>>>(dst[k], s->pp_tmp, taps, 0, g1);
        gaussian_pass_kernel<<<grd, blk, 0, s->stream>>>(s->pp_tmp, dst[k], taps, 1, g1);
    }
    int rc;
    if (fam == FAM_TV || fam == FAM_NLTV) {
        float *set0 = s->state;
        if ((rc = up2d(s, set0 + (ST_U1 * B + slot) * P, u))) return rc;
        if ((rc = up2d(s, set0 + (ST_U2 * B + slot) * P, u + n))) return rc;
        for (int k = ST_XI11; k <= ST_XI22; k++)
            FALDOI_CUDA(cudaMemsetAsync(set0 + (k * B + slot) * P, 0, P * sizeof(float), s->stream));
        FALDOI_CUDA(cudaMemsetAsync(s->parity + slot, 0, sizeof(int), s->stream));
    }
    if (fam == FAM_NLTV) {
        image_to_lab_kernel<<<grd, blk, 0, s->stream>>>(s->raw_stage, s->lab + slot * P, s->lab + (B + slot) * P,
                                                        s->lab + (2 * B + slot) * P, g1);
        for (int k = 0; k < 2 * NL_SLOTS; k++)
            FALDOI_CUDA(cudaMemsetAsync(s->dual + (k * B + slot) * P, 0, P * sizeof(float), s->stream));
        s->nl_parity[slot] = 0;
    }
    if (fam == FAM_OCC) {
        if ((rc = up2d(s, s->occ + (OC_U1 * B + slot) * P, u))) return rc;
        if ((rc = up2d(s, s->occ + (OC_U2 * B + slot) * P, u + n))) return rc;
        if ((rc = up2d(s, s->occ + (OC_CHI0 * B + slot) * P, chi))) return rc;
        FALDOI_CUDA(cudaMemsetAsync(s->occ + (OC_ETA0 * B + slot) * P, 0, P * sizeof(float), s->stream));
        FALDOI_CUDA(cudaMemsetAsync(s->occ + ((OC_ETA0 + 1) * B + slot) * P, 0, P * sizeof(float), s->stream));
    }
    FALDOI_CUDA(cudaGetLastError());
    return FALDOI_OK;
}

// test / inspection hook: the preprocessed frames of a slot back to the host (dense w*h each; any may be NULL)
extern "C" int faldoi_solver_download_frames(faldoi_solver *s, int slot, float *I0n, float *I1n, float *Im1n, float *lab) {
    if (!s || slot < 0 || slot >= s->B) return FALDOI_ERR_ARG;
    FALDOI_CUDA(cudaSetDevice(s->device));
    const Geo g = s->g;
    const size_t P = g.plane, B = s->B, n = (size_t)g.w * g.h;
    auto dn = [&](float *dst, const float *src) {
        return cudaMemcpy2DAsync(dst, g.w * sizeof(float), src, g.pitch * sizeof(float), g.w * sizeof(float), g.h,
                                 cudaMemcpyDeviceToHost, s->stream);
    };
    if (I0n) FALDOI_CUDA(dn(I0n, s->I0 + slot * P));
    if (I1n) FALDOI_CUDA(dn(I1n, s->I1 + slot * P));
    if (Im1n && s->occ) FALDOI_CUDA(dn(Im1n, s->occ + (OC_IM1 * B + slot) * P));
    if (lab && s->lab)
        for (int c = 0; c < 3; c++) FALDOI_CUDA(dn(lab + c * n, s->lab + (c * B + slot) * P));
    FALDOI_CUDA(cudaStreamSynchronize(s->stream));
    return FALDOI_OK;
}

// ---------------------------------------------------------------------------
// end-of-warp bookkeeping: iterations run, last error, ping-pong parity
// ---------------------------------------------------------------------------
__global__ void finalize_warp_kernel(const unsigned *err_max, const double *err_sum, int use_sum, int always_all,
                                     int *parity, int *log_iters, float *log_err, int max_iters, float tol2,
                                     float npix, int warp_idx, int npairs, int iters_per_flip = 1) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= npairs) return;
    int n = 0;
    float e = INFINITY;
    while (n < max_iters) {
        e = use_sum ? (float)err_sum[(size_t)b * max_iters + n] / npix : __uint_as_float(err_max[(size_t)b * max_iters + n]);
        n++;
        if (!always_all && !(e > tol2)) break;
    }
    // ping-pong flips: one per iteration, or one per launch of the two-iteration kernel (ceil(n/2))
    if (parity) parity[b] = (parity[b] + (n + iters_per_flip - 1) / iters_per_flip) & 1;
    log_iters[b * FALDOI_MAX_WARPS + warp_idx] = n;
    log_err[b * FALDOI_MAX_WARPS + warp_idx] = e;
}

// number of pairs whose exit test still holds after iteration `it` (see pair_active)
__global__ void count_active_kernel(const unsigned *err_max, const double *err_sum, int use_sum, int npairs, int max_iters,
                                    int it, float tol2, float npix, int *out) {
    int n = 0;
    for (int b = 0; b < npairs; b++) {
        const float e = use_sum ? (float)err_sum[(size_t)b * max_iters + it] / npix : __uint_as_float(err_max[(size_t)b * max_iters + it]);
        n += (e > tol2);
    }
    *out = n;
}

// same for the two-iteration kernel after its launch L: a pair keeps iterating iff that launch ran
// normally and both of its iterations still exceed the tolerance
__global__ void count_active2_kernel(const unsigned *err_max, const unsigned char *stat, int stat_stride, int npairs, int max_iters,
                                     int L, float tol2, int *out) {
    int n = 0;
    for (int b = 0; b < npairs; b++) {
        const unsigned *e = err_max + (size_t)b * max_iters + 2 * L;
        bool act = stat[(size_t)b * stat_stride + L] && __uint_as_float(e[0]) > tol2;
        if (act && 2 * L + 1 < max_iters) act = __uint_as_float(e[1]) > tol2;
        n += act;
    }
    *out = n;
}

// the same with TV-CSAD's mean error (tv_csad2_kernel)
__global__ void count_active2_sum_kernel(const double *err_sum, const unsigned char *stat, int stat_stride, int npairs, int max_iters,
                                         int L, float tol2, float npix, int *out) {
    int n = 0;
    for (int b = 0; b < npairs; b++) {
        const double *e = err_sum + (size_t)b * max_iters + 2 * L;
        bool act = stat[(size_t)b * stat_stride + L] && (float)e[0] / npix > tol2;
        if (act && 2 * L + 1 < max_iters) act = (float)e[1] / npix > tol2;
        n += act;
    }
    *out = n;
}

__global__ void export_flow_kernel(const float *state, size_t set_stride, const int *parity, float *packed, Geo g) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= g.w || y >= g.h) return;
    const float *set = state + (size_t)parity[b] * set_stride + (size_t)b * g.plane;
    const size_t ks = (size_t)g.B * g.plane, n = (size_t)g.w * g.h;
    float *o = packed + (size_t)b * 3 * n;
    o[(size_t)y * g.w + x] = set[ST_U1 * ks + (size_t)y * g.pitch + x];
    o[n + (size_t)y * g.w + x] = set[ST_U2 * ks + (size_t)y * g.pitch + x];
}

static dim3 grid2d(const Geo &g, dim3 block, int npairs) {
    return dim3((g.w + block.x - 1) / block.x, (g.h + block.y - 1) / block.y, npairs);
}

// DivConst for x/theta: exhaustively verified on the device (2^23 significands) the first
// time a theta value is seen by this handle; a failed verification just selects IEEE division.
static int make_div_const(faldoi_solver *s, float b, DivConst *out) {
    if (s->dc_valid && s->dc.b == b) {
        *out = s->dc;
        return FALDOI_OK;
    }
    DivConst d{b, 1.0f / b, 0};
    if (!s->dc_flag) {
        s->dc_flag = (int *)s->dmalloc(1);
        if (!s->dc_flag) return FALDOI_ERR_MEM;
    }
    FALDOI_CUDA(cudaMemsetAsync(s->dc_flag, 0, sizeof(int), s->stream));
    verify_div_const_kernel<<<(1u << 23) / 256, 256, 0, s->stream>>>(d.b, d.rb, s->dc_flag);
    int mismatch = 1;
    FALDOI_CUDA(cudaMemcpyAsync(&mismatch, s->dc_flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    FALDOI_CUDA(cudaStreamSynchronize(s->stream));
    d.ok = (mismatch == 0) ? 1 : 0;
    s->dc = d;
    s->dc_valid = true;
    *out = d;
    return FALDOI_OK;
}

static int run_tv(faldoi_solver *s, const faldoi_params *p, int npairs) {
    const Geo g = s->g;
    const bool csad = method_is_csad(s->method);
    const dim3 blk(32, 8);
    centered_gradient_kernel<<<grid2d(g, blk, npairs), blk, 0, s->stream>>>(s->I1, s->I1x, s->I1y, g);
    s->launches++;
    const size_t ks = (size_t)g.B * g.plane;
    TvArgs a{};
    a.state = s->state;
    a.set_stride = s->set_stride;
    a.Ix = s->Ix;
    a.Iy = s->Iy;
    a.rho_c = s->rho_c;
    a.scale = s->scale;
    a.err_max = s->err_max;
    a.err_chk = s->err_max;
    a.err_sum = s->err_sum;
    a.parity = s->parity;
    a.g = g;
    a.max_iters = p->max_iters;
    a.tau = p->tau;
    a.theta = p->theta;
    a.l_t = p->lambda * p->theta;
    a.tol2 = p->tol * p->tol;
    if (int rc = make_div_const(s, p->theta, &a.dth)) return rc;
    // Both models run two iterations per HBM pass (tv_tile2_kernel / tv_csad2_kernel); stat[b][L] = launch L ran normally
    {
        const int need = p->max_iters / 2 + 4;
        if (need > s->t2_stride) {
            s->t2_stat = (unsigned char *)s->dmalloc(((size_t)g.B * need + 3) / 4);
            if (csad) s->csad_ticket = (unsigned *)s->dmalloc((size_t)g.B * need);
            if (!s->t2_stat || (csad && !s->csad_ticket)) return FALDOI_ERR_MEM;
            s->t2_stride = need;
        }
    }
    for (int wp = 0; wp < p->warps; wp++) {
        if (csad) FALDOI_CUDA(cudaMemsetAsync(s->csad_ticket, 0, (size_t)g.B * s->t2_stride * sizeof(unsigned), s->stream));
        if (csad)
            FALDOI_CUDA(cudaMemsetAsync(s->err_sum, 0, (size_t)g.B * p->max_iters * sizeof(double), s->stream));
        else
            FALDOI_CUDA(cudaMemsetAsync(s->err_max, 0, (size_t)g.B * p->max_iters * sizeof(unsigned), s->stream));
        WarpArgs wa{};
        wa.I0 = s->I0;
        wa.I1 = s->I1;
        wa.I1x = s->I1x;
        wa.I1y = s->I1y;
        wa.i1_plane = s->i1_plane;
        wa.u1 = s->state + ST_U1 * ks;
        wa.u2 = s->state + ST_U2 * ks;
        wa.ub1 = s->state + ST_UB1 * ks;
        wa.ub2 = s->state + ST_UB2 * ks;
        wa.Ix = s->Ix;
        wa.Iy = s->Iy;
        wa.rho_c = csad ? nullptr : s->rho_c;
        wa.I1w = csad ? s->I1w : nullptr;
        wa.parity = s->parity;
        wa.set_stride = s->set_stride;
        wa.g = g;
        warp_constants_kernel<<<grid2d(g, blk, npairs), blk, 0, s->stream>>>(wa);
        s->launches++;
        if (csad) {
            CsadPermArgs ca{};
            ca.I0 = s->I0;
            ca.I1w = s->I1w;
            ca.Ix = s->Ix;
            ca.Iy = s->Iy;
            ca.u1 = wa.u1;
            ca.u2 = wa.u2;
            ca.parity = s->parity;
            ca.set_stride = s->set_stride;
            ca.scale = s->scale;
            ca.t1 = s->csad_t1;
            ca.t2 = s->csad_t2;
            ca.perm = s->csad_perm;
            ca.g = g;
            const dim3 cb(32, 4);
            csad_perm_kernel<<<grid2d(g, cb, npairs), cb, 0, s->stream>>>(ca);
            s->launches++;
        }
        if (s->phase_mark()) return FALDOI_ERR_CUDA;
        // All launches of a warp are enqueued without waiting for results; converged pairs return at
        // once.  To stop enqueuing once EVERY pair has met the exit test, the active-pair count is
        // copied to pinned memory every LCHUNK launches and inspected two chunks later (so the GPU
        // always has at least one full chunk queued and never waits for the host).
        // Two iterations per launch: ceil(max_iters/2) launches + one slot for a trailing fix-up.
        const int nL = (p->max_iters + 1) / 2 + 1, LCHUNK = 13;
        FALDOI_CUDA(cudaMemsetAsync(s->t2_stat, 0, (size_t)g.B * s->t2_stride, s->stream));
        const dim3 grid((g.pitch + T2_W - 1) / T2_W, (g.h + T2_H - 1) / T2_H, npairs);
        const dim3 cgrid((g.pitch + C2_W - 1) / C2_W, (g.h + C2_H - 1) / C2_H, npairs);
        const T2Args t2{s->t2_stat, s->t2_stride};
        const Csad2Args c2{s->csad_perm, s->t2_stat, s->t2_stride, s->csad_partial, s->csad_ticket};
        for (int c = 0, L = 0; L < nL; c++) {
            if (c >= 2) {
                FALDOI_CUDA(cudaEventSynchronize(s->chunk_ev[(c - 2) & 3]));
                if (s->h_active[(c - 2) & 3] == 0) break;
            }
            const int end = (L + LCHUNK < nL) ? L + LCHUNK : nL;
            for (; L < end; L++) {
                if (csad)
                    tv_csad2_kernel<<<cgrid, C2_THREADS, sizeof(Csad2Smem) + 128, s->stream>>>(s->cmaps, a, c2, L);
                else
                    tv_tile2_kernel<<<grid, T2_THREADS, sizeof(Tile2Smem) + 128, s->stream>>>(s->maps2, a, t2, L);
                s->launches++;
            }
            const int Lc = (L - 1 < (p->max_iters + 1) / 2) ? L - 1 : (p->max_iters + 1) / 2 - 1;
            if (csad)
                count_active2_sum_kernel<<<1, 1, 0, s->stream>>>(s->err_sum, s->t2_stat, s->t2_stride, npairs, p->max_iters, Lc, a.tol2,
                                                                 (float)(g.w * g.h), s->d_active + (c & 3));
            else
                count_active2_kernel<<<1, 1, 0, s->stream>>>(s->err_max, s->t2_stat, s->t2_stride, npairs, p->max_iters, Lc, a.tol2,
                                                             s->d_active + (c & 3));
            FALDOI_CUDA(cudaMemcpyAsync(s->h_active + (c & 3), s->d_active + (c & 3), sizeof(int), cudaMemcpyDeviceToHost, s->stream));
            FALDOI_CUDA(cudaEventRecord(s->chunk_ev[c & 3], s->stream));
            s->launches++;
        }
        if (s->phase_mark()) return FALDOI_ERR_CUDA;
        finalize_warp_kernel<<<(npairs + 63) / 64, 64, 0, s->stream>>>(s->err_max, s->err_sum, csad ? 1 : 0, 0, s->parity,
                                                                       s->log_iters, s->log_err, p->max_iters, a.tol2,
                                                                       (float)(g.w * g.h), wp, npairs, 2);
        s->launches++;
    }
    export_flow_kernel<<<grid2d(g, blk, npairs), blk, 0, s->stream>>>(s->state, s->set_stride, s->parity, s->packed, g);
    s->launches++;
    FALDOI_CUDA(cudaGetLastError());
    return FALDOI_OK;
}


// ---------------------------------------------------------------------------
// NLTV family (methods 2,3,6,7)
// ---------------------------------------------------------------------------
static int run_nltv(faldoi_solver *s, const faldoi_params *p, int npairs) {
    const Geo g = s->g;
    const bool csad = method_is_csad(s->method);
    const dim3 blk(32, 8);
    const size_t ks = (size_t)g.B * g.plane;
    NlOffsets offs;
    for (int sl = 0; sl < NL_SLOTS; sl++) {
        int k, l;
        nl_slot_offset(sl, k, l);
        // get_wspatial_2 (src/global_faldoi.cpp:943-952): difS = hypot(l,k) (double) stored as float
        const float difS = (float)hypot((double)l, (double)k);
        offs.ws[sl] = expf(-difS / 2.f);
    }
    nltv_init_kernel<<<grid2d(g, blk, npairs), blk, 0, s->stream>>>(s->lab, s->wgt, s->wt, s->rwt, offs, g);
    centered_gradient_kernel<<<grid2d(g, blk, npairs), blk, 0, s->stream>>>(s->I1, s->I1x, s->I1y, g);
    s->launches += 2;
    NlArgs a{};
    a.state = s->state;
    a.set_stride = s->set_stride;
    a.dual = s->dual;
    a.dual_set_stride = s->dual_set_stride;
    a.wgt = s->wgt;
    a.wt = s->wt;
    a.Ix = s->Ix;
    a.Iy = s->Iy;
    a.rho_c = s->rho_c;
    a.scale = s->scale;
    a.blk = s->csad_blk;
    a.sep = s->csad_sep;
    a.err_sum = s->err_sum;
    a.g = g;
    a.max_iters = p->max_iters;
    a.tau = p->tau;
    a.theta = p->theta;
    a.l_t = p->lambda * p->theta;
    if (int rc = make_div_const(s, p->theta, &a.dth)) return rc;
    // Every pair runs all max_iters iterations, so one ping-pong parity serves the whole batch; it is carried from
    // run to run (a second run without a fresh upload continues from the first one's result, as the TV family does).
    // Slots uploaded at different times would disagree -- refuse that instead of reading a stale set.
    int base_parity = s->nl_parity[0];
    for (int b = 1; b < npairs; b++)
        if (s->nl_parity[b] != base_parity) {
            set_error("faldoi_solver_run (NLTV): the slots of a batch must have been uploaded together (they carry different run histories)");
            return FALDOI_ERR_ARG;
        }
    for (int wp = 0; wp < p->warps; wp++) {
        FALDOI_CUDA(cudaMemsetAsync(s->err_sum, 0, (size_t)g.B * p->max_iters * sizeof(double), s->stream));
        WarpArgs wa{};
        wa.I0 = s->I0;
        wa.I1 = s->I1;
        wa.I1x = s->I1x;
        wa.I1y = s->I1y;
        wa.i1_plane = s->i1_plane;
        wa.u1 = s->state + ST_U1 * ks;
        wa.u2 = s->state + ST_U2 * ks;
        wa.ub1 = s->state + ST_UB1 * ks;
        wa.ub2 = s->state + ST_UB2 * ks;
        wa.Ix = s->Ix;
        wa.Iy = s->Iy;
        wa.rho_c = csad ? nullptr : s->rho_c;
        wa.I1w = csad ? s->I1w : nullptr;
        wa.parity = s->parity;
        wa.set_stride = s->set_stride;
        wa.g = g;
        warp_constants_kernel<<<grid2d(g, blk, npairs), blk, 0, s->stream>>>(wa);
        s->launches++;
        if (csad) {
            CsadArgs ca{};
            ca.I0 = s->I0;
            ca.I1w = s->I1w;
            ca.Ix = s->Ix;
            ca.Iy = s->Iy;
            ca.u1 = wa.u1;
            ca.u2 = wa.u2;
            ca.parity = s->parity;
            ca.set_stride = s->set_stride;
            ca.scale = s->scale;
            ca.blk = s->csad_blk;
            ca.sep = s->csad_sep;
            ca.g = g;
            ca.hyp = 0;
            const dim3 cb(32, 4);
            csad_constants_kernel<<<grid2d(g, cb, npairs), cb, 0, s->stream>>>(ca);
            s->launches++;
        }
        if (s->phase_mark()) return FALDOI_ERR_CUDA;
        // Exact arithmetic (bit-identical to the reference) unless the caller opted into the approximate mode
        // (faldoi_solver_set_nltv_fast: approximate divisions, paired slot order).
        const bool fast = s->nltv_fast;
        const dim3 tgrid((g.pitch + NLT_W - 1) / NLT_W, (g.h + NLT_H - 1) / NLT_H, npairs);
        const size_t tsm = sizeof(NlTileSmem) + 128;
        for (int it = 0; it < p->max_iters; it++) {
            if (fast) {
                if (csad)
                    nltv_tile_kernel<DATA_CSAD, false><<<tgrid, NLT_THREADS, tsm, s->stream>>>(s->nlmaps, a, it, base_parity);
                else
                    nltv_tile_kernel<DATA_TVL1, false><<<tgrid, NLT_THREADS, tsm, s->stream>>>(s->nlmaps, a, it, base_parity);
            } else {
                if (csad)
                    nltv_tile_kernel<DATA_CSAD, true><<<tgrid, NLT_THREADS, tsm, s->stream>>>(s->nlmaps, a, it, base_parity);
                else
                    nltv_tile_kernel<DATA_TVL1, true><<<tgrid, NLT_THREADS, tsm, s->stream>>>(s->nlmaps, a, it, base_parity);
            }
        }
        if (s->phase_mark()) return FALDOI_ERR_CUDA;
        s->launches += p->max_iters;
        finalize_warp_kernel<<<(npairs + 63) / 64, 64, 0, s->stream>>>(s->err_max, s->err_sum, 1, 1, s->parity, s->log_iters,
                                                                       s->log_err, p->max_iters, 0.f, (float)(g.w * g.h),
                                                                       wp, npairs);
        s->launches++;
        base_parity = (base_parity + p->max_iters) & 1;
    }
    for (int b = 0; b < npairs; b++) s->nl_parity[b] = base_parity;
    export_flow_kernel<<<grid2d(g, blk, npairs), blk, 0, s->stream>>>(s->state, s->set_stride, s->parity, s->packed, g);
    s->launches++;
    FALDOI_CUDA(cudaGetLastError());
    return FALDOI_OK;
}

// ---------------------------------------------------------------------------
// TVL2 + occlusions (method 8)
// ---------------------------------------------------------------------------
static int occ_upload(faldoi_solver *s, int slot, const float *Im1, const float *u, const float *chi) {
    const size_t P = s->g.plane, B = s->B, n = (size_t)s->g.w * s->g.h;
    int rc;
    if ((rc = up2d(s, s->occ + (OC_IM1 * B + slot) * P, Im1))) return rc;
    if ((rc = up2d(s, s->occ + (OC_U1 * B + slot) * P, u))) return rc;
    if ((rc = up2d(s, s->occ + (OC_U2 * B + slot) * P, u + n))) return rc;
    if ((rc = up2d(s, s->occ + (OC_CHI0 * B + slot) * P, chi))) return rc;
    // eta = 0 at entry (target definition; the reference never initialises it)
    FALDOI_CUDA(cudaMemsetAsync(s->occ + (OC_ETA0 * B + slot) * P, 0, P * sizeof(float), s->stream));
    FALDOI_CUDA(cudaMemsetAsync(s->occ + ((OC_ETA0 + 1) * B + slot) * P, 0, P * sizeof(float), s->stream));
    return FALDOI_OK;
}

// sweeps fused per launch by the temporally blocked OCC kernels: 24/NS launches, which must be an
// even number so the ping-pong ends in set 0 where the u-update / next outer iteration read
#ifndef FALDOI_OCC_NS
#define FALDOI_OCC_NS 4
#endif
enum { OCC_NS = FALDOI_OCC_NS };
static_assert(24 % OCC_NS == 0 && (24 / OCC_NS) % 2 == 0, "24/OCC_NS must be an even integer");

static int run_occ(faldoi_solver *s, const faldoi_params *p, int npairs) {
    const Geo g = s->g;
    const dim3 blk(32, 8);
    const dim3 grd = grid2d(g, blk, npairs);
    static_assert(OCC_NS <= 4, "the register-resident OCC kernels stage a 4-column apron");
    const dim3 rgrd((g.pitch + OR_W - 1) / OR_W, (g.h + OR_TH - 1) / OR_TH, npairs);
    OccArgs a{};
    a.pl = s->occ;
    a.I0 = s->I0;
    a.I1 = s->I1;
    a.I1x = s->I1x;
    a.I1y = s->I1y;
    a.err_max = s->err_max;
    a.g = g;
    a.max_iters = p->max_iters;
    a.lambda = p->lambda;
    a.theta = p->theta;
    a.beta = p->beta;
    a.alpha = p->alpha;
    a.tau_theta = p->tau_u / p->theta;
    a.mu = p->mu;
    a.tau_eta = p->tau_eta;
    a.tau_chi = p->tau_chi;
    a.l_t = p->lambda * p->theta;
    a.tol2 = p->tol * p->tol;
    centered_gradient_kernel<<<grd, blk, 0, s->stream>>>(s->I1, s->I1x, s->I1y, g);
    occ_init_kernel<<<grd, blk, 0, s->stream>>>(a);
    s->launches += 2;
    for (int wp = 0; wp < p->warps; wp++) {
        FALDOI_CUDA(cudaMemsetAsync(s->err_max, 0, (size_t)g.B * p->max_iters * sizeof(unsigned), s->stream));
        occ_warp_kernel<<<grd, blk, 0, s->stream>>>(a);
        s->launches++;
        if (s->phase_mark()) return FALDOI_ERR_CUDA;
        for (int it = 0; it < p->max_iters; it++) {
            occ_v_kernel<<<grd, blk, 0, s->stream>>>(a, it);
            for (int k = 0; k < 24 / OCC_NS; k++)
                occ_xi_rows_kernel<OCC_NS><<<rgrd, 32 * (OR_TH + 2 * OCC_NS), 0, s->stream>>>(a, it, k & 1);
            occ_u_kernel<<<grd, blk, 0, s->stream>>>(a, it);
            for (int k = 0; k < 24 / OCC_NS; k++)
                occ_chi_rows_kernel<OCC_NS><<<rgrd, 32 * (OR_TH + 2 * OCC_NS), 0, s->stream>>>(a, it, k & 1, k == 24 / OCC_NS - 1);
            s->launches += 2 + 2 * (24 / OCC_NS);
        }
        if (s->phase_mark()) return FALDOI_ERR_CUDA;
        finalize_warp_kernel<<<(npairs + 63) / 64, 64, 0, s->stream>>>(s->err_max, s->err_sum, 0, 0, nullptr, s->log_iters,
                                                                       s->log_err, p->max_iters, a.tol2, (float)(g.w * g.h),
                                                                       wp, npairs);
        s->launches++;
    }
    occ_export_kernel<<<grd, blk, 0, s->stream>>>(a, s->packed);
    s->launches++;
    FALDOI_CUDA(cudaGetLastError());
    return FALDOI_OK;
}

extern "C" int faldoi_solver_run(faldoi_solver *s, const faldoi_params *p, int npairs) {
    if (!s || !p || npairs < 1 || npairs > s->B || p->warps < 0 || p->warps > FALDOI_MAX_WARPS || p->max_iters < 0) {
        set_error("faldoi_solver_run: bad argument");
        return FALDOI_ERR_ARG;
    }
    if (method_family(p->method) != method_family(s->method) || method_is_csad(p->method) != method_is_csad(s->method)) {
        set_error("faldoi_solver_run: params.method does not match the handle's method family");
        return FALDOI_ERR_ARG;
    }
    FALDOI_CUDA(cudaSetDevice(s->device));
    if (s->alloc_err(p->max_iters > 0 ? p->max_iters : 1) != FALDOI_OK) return FALDOI_ERR_MEM;
    s->launches = 0;
    s->phase_used = 0;
    FALDOI_CUDA(cudaMemsetAsync(s->log_iters, 0, (size_t)s->B * FALDOI_MAX_WARPS * sizeof(int), s->stream));
    FALDOI_CUDA(cudaEventRecord(s->ev0, s->stream));
    int rc;
    switch (method_family(s->method)) {
        case FAM_TV: rc = run_tv(s, p, npairs); break;
        case FAM_NLTV: rc = run_nltv(s, p, npairs); break;
        default: rc = run_occ(s, p, npairs); break;
    }
    if (rc != FALDOI_OK) return rc;
    FALDOI_CUDA(cudaEventRecord(s->ev1, s->stream));
    s->ran = true;
    return FALDOI_OK;
}

extern "C" int faldoi_solver_sync(faldoi_solver *s) {
    if (!s) return FALDOI_ERR_ARG;
    FALDOI_CUDA(cudaSetDevice(s->device));
    FALDOI_CUDA(cudaStreamSynchronize(s->stream));
    if (s->ran) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess) s->last_ms = ms;
        float acc = 0.f;
        for (int i = 0; i + 1 < s->phase_used; i += 2)
            if (cudaEventElapsedTime(&ms, s->phase_ev[i], s->phase_ev[i + 1]) == cudaSuccess) acc += ms;
        s->last_iter_ms = acc;
    }
    return FALDOI_OK;
}

extern "C" int faldoi_solver_download(faldoi_solver *s, int slot, float *u, float *chi, faldoi_log *log) {
    if (!s || slot < 0 || slot >= s->B || !u) {
        set_error("faldoi_solver_download: bad argument");
        return FALDOI_ERR_ARG;
    }
    FALDOI_CUDA(cudaSetDevice(s->device));
    const size_t n = (size_t)s->g.w * s->g.h;
    FALDOI_CUDA(cudaMemcpyAsync(u, s->packed + (size_t)slot * 3 * n, 2 * n * sizeof(float), cudaMemcpyDefault, s->stream));
    if (chi && s->method == FALDOI_M_TVL1_OCC)
        FALDOI_CUDA(cudaMemcpyAsync(chi, s->packed + (size_t)slot * 3 * n + 2 * n, n * sizeof(float), cudaMemcpyDefault,
                                    s->stream));
    if (log) {
        FALDOI_CUDA(cudaMemcpyAsync(log->iters, s->log_iters + slot * FALDOI_MAX_WARPS, sizeof(log->iters),
                                    cudaMemcpyDeviceToHost, s->stream));
        FALDOI_CUDA(cudaMemcpyAsync(log->err, s->log_err + slot * FALDOI_MAX_WARPS, sizeof(log->err),
                                    cudaMemcpyDeviceToHost, s->stream));
    }
    return faldoi_solver_sync(s);
}

// Initial / final dual variables of the TV family (methods 0,1,4,5): the xi11..xi22 arrays of tvl2OF / tvcsad_PD.
extern "C" int faldoi_solver_upload_xi(faldoi_solver *s, int slot, const float *xi11, const float *xi12, const float *xi21,
                                       const float *xi22) {
    if (!s || slot < 0 || slot >= s->B || !xi11 || !xi12 || !xi21 || !xi22 || method_family(s->method) != FAM_TV) {
        set_error("faldoi_solver_upload_xi: bad argument (TV-family handles only)");
        return FALDOI_ERR_ARG;
    }
    FALDOI_CUDA(cudaSetDevice(s->device));
    const float *src[4] = {xi11, xi12, xi21, xi22};
    // right after an upload the slot's parity is 0: the initial state lives in set 0
    for (int k = 0; k < 4; k++)
        if (int rc = up2d(s, s->state + ((size_t)(ST_XI11 + k) * s->B + slot) * s->g.plane, src[k])) return rc;
    return FALDOI_OK;
}

extern "C" int faldoi_solver_download_xi(faldoi_solver *s, int slot, float *xi11, float *xi12, float *xi21, float *xi22) {
    if (!s || slot < 0 || slot >= s->B || !xi11 || !xi12 || !xi21 || !xi22 || method_family(s->method) != FAM_TV) {
        set_error("faldoi_solver_download_xi: bad argument (TV-family handles only)");
        return FALDOI_ERR_ARG;
    }
    FALDOI_CUDA(cudaSetDevice(s->device));
    int par = 0;
    FALDOI_CUDA(cudaMemcpyAsync(&par, s->parity + slot, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    FALDOI_CUDA(cudaStreamSynchronize(s->stream));
    float *dst[4] = {xi11, xi12, xi21, xi22};
    const Geo &g = s->g;
    for (int k = 0; k < 4; k++)
        FALDOI_CUDA(cudaMemcpy2DAsync(dst[k], g.w * sizeof(float), s->state + (size_t)par * s->set_stride + ((size_t)(ST_XI11 + k) * s->B + slot) * g.plane,
                                      g.pitch * sizeof(float), g.w * sizeof(float), g.h, cudaMemcpyDeviceToHost, s->stream));
    FALDOI_CUDA(cudaStreamSynchronize(s->stream));
    return FALDOI_OK;
}

// Page-locked host memory for staging (H2D / D2H copies from it are truly asynchronous and run at full PCIe rate).
extern "C" void *faldoi_pinned_alloc(size_t bytes) {
    void *p = nullptr;
    if (!cuda_ok(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable), "cudaHostAlloc")) return nullptr;
    return p;
}
extern "C" void faldoi_pinned_free(void *p) {
    if (p) cudaFreeHost(p);
}

extern "C" int faldoi_solver_set_nltv_fast(faldoi_solver *s, int fast) {
    if (!s || method_family(s->method) != FAM_NLTV) {
        set_error("faldoi_solver_set_nltv_fast: not an NLTV handle");
        return FALDOI_ERR_ARG;
    }
    s->nltv_fast = fast != 0;
    return FALDOI_OK;
}

extern "C" float faldoi_solver_last_run_ms(faldoi_solver *s) { return s ? s->last_ms : -1.f; }
extern "C" float faldoi_solver_last_iter_ms(faldoi_solver *s) { return s ? s->last_iter_ms : -1.f; }
extern "C" long long faldoi_solver_last_launches(faldoi_solver *s) { return s ? s->launches : -1; }
extern "C" void *faldoi_solver_stream(faldoi_solver *s) { return s ? (void *)s->stream : nullptr; }
extern "C" float *faldoi_solver_device_flow(faldoi_solver *s, int slot, int *pitch_floats) {
    if (!s || slot < 0 || slot >= s->B) return nullptr;
    if (pitch_floats) *pitch_floats = s->g.w;
    return s->packed + (size_t)slot * 3 * s->g.w * s->g.h;
}

// ---------------------------------------------------------------------------
// one-call host entry and the per-solver mirrors
// ---------------------------------------------------------------------------
namespace {
// Single-pair handles behind the one-call entries, kept alive between calls so that a sequence of pairs does not
// pay cudaMalloc per call.  The cache is PER HOST THREAD (a few entries, least recently used first out): threads
// driving different GPUs -- or the same one -- never share a handle, so the entries need no lock and eight host
// threads on eight GPUs run concurrently (re-entrant per device).
struct SolverCache {
    struct Entry {
        int device, w, h, method;
        faldoi_solver *s;
    };
    std::vector<Entry> e;
    ~SolverCache() {
        for (Entry &x : e) faldoi_solver_destroy(x.s);
    }
    int get(int device, int w, int h, int method, faldoi_solver **out) {
        for (size_t i = 0; i < e.size(); i++)
            if (e[i].device == device && e[i].w == w && e[i].h == h && e[i].method == method) {
                const Entry hit = e[i];
                e.erase(e.begin() + i);
                e.push_back(hit);
                *out = hit.s;
                return FALDOI_OK;
            }
        if (e.size() >= 3) {
            faldoi_solver_destroy(e.front().s);
            e.erase(e.begin());
        }
        faldoi_solver *s = nullptr;
        const int rc = faldoi_solver_create(&s, device, w, h, method, 1);
        if (rc != FALDOI_OK) return rc;
        e.push_back(Entry{device, w, h, method, s});
        *out = s;
        return FALDOI_OK;
    }
};
thread_local SolverCache t_cache;
}  // namespace

extern "C" int faldoi_global_solve(int device, const faldoi_params *p, int w, int h, const float *I0, const float *I1,
                                   const float *Im1, const float *lab, float *u, float *chi, faldoi_log *log) {
    if (!p || !I0 || !I1 || !u) {
        set_error("faldoi_global_solve: null argument");
        return FALDOI_ERR_ARG;
    }
    faldoi_solver *s = nullptr;
    int rc = t_cache.get(device, w, h, p->method, &s);
    if (rc != FALDOI_OK) return rc;
    if ((rc = faldoi_solver_upload(s, 0, I0, I1, Im1, lab, u, chi)) != FALDOI_OK) return rc;
    if ((rc = faldoi_solver_run(s, p, 1)) != FALDOI_OK) return rc;
    return faldoi_solver_download(s, 0, u, chi, log);
}

extern "C" int faldoi_global_solve_raw(int device, const faldoi_params *p, int w, int h, int pd, const float *i0, const float *i1,
                                       const float *im1, float *u, float *chi, faldoi_log *log) {
    if (!p || !i0 || !i1 || !im1 || !u) {
        set_error("faldoi_global_solve_raw: null argument");
        return FALDOI_ERR_ARG;
    }
    faldoi_solver *s = nullptr;
    int rc = t_cache.get(device, w, h, p->method, &s);
    if (rc != FALDOI_OK) return rc;
    if ((rc = faldoi_solver_upload_raw(s, 0, i0, i1, im1, pd, u, chi)) != FALDOI_OK) return rc;
    if ((rc = faldoi_solver_run(s, p, 1)) != FALDOI_OK) return rc;
    return faldoi_solver_download(s, 0, u, chi, log);
}

static void print_log(const faldoi_params &p, const faldoi_log &log, FILE *f, const char *fmt) {
    for (int k = 0; k < p.warps; k++) fprintf(f, fmt, k, log.iters[k], log.err[k]);
}

static int solve_split(const faldoi_params &p, int nx, int ny, const float *I0, const float *I1, const float *Im1,
                       const float *lab, float *u1, float *u2, float *chi, float *const xi[4], int verbose, FILE *vf, const char *fmt) {
    const size_t n = (size_t)nx * ny;
    std::vector<float> u(2 * n);
    memcpy(u.data(), u1, n * sizeof(float));
    memcpy(u.data() + n, u2, n * sizeof(float));
    faldoi_log log{};
    faldoi_solver *s = nullptr;
    int rc = t_cache.get(0, nx, ny, p.method, &s);
    if (rc != FALDOI_OK) return rc;
    if ((rc = faldoi_solver_upload(s, 0, I0, I1, Im1, lab, u.data(), chi)) != FALDOI_OK) return rc;
    const bool with_xi = xi && xi[0] && xi[1] && xi[2] && xi[3];
    if (with_xi && (rc = faldoi_solver_upload_xi(s, 0, xi[0], xi[1], xi[2], xi[3])) != FALDOI_OK) return rc;
    if ((rc = faldoi_solver_run(s, &p, 1)) != FALDOI_OK) return rc;
    if ((rc = faldoi_solver_download(s, 0, u.data(), chi, &log)) != FALDOI_OK) return rc;
    if (with_xi && (rc = faldoi_solver_download_xi(s, 0, xi[0], xi[1], xi[2], xi[3])) != FALDOI_OK) return rc;
    memcpy(u1, u.data(), n * sizeof(float));
    memcpy(u2, u.data() + n, n * sizeof(float));
    if (verbose) print_log(p, log, vf, fmt);
    return FALDOI_OK;
}

// The reference's xi arguments are caller-owned arrays that the solver reads as the initial dual variables and
// updates in place (src/global_faldoi.cpp:556-573, 735; main() zeroes them right before the call, :2116-2121).
// The mirrors honour that: the four arrays are uploaded as the initial duals and hold the final duals on return
// (a warm-starting caller sees what it would see from the reference).  Passing NULL for them starts from zero.
extern "C" int faldoi_tvl2OF(const float *I0, float *I1, float *u1, float *u2, float *xi11, float *xi12, float *xi21, float *xi22,
                             float lambda, float theta, float tau, float tol_OF, int nx, int ny, int warps,
                             int verbose) {
    faldoi_params p;
    faldoi_default_params(FALDOI_M_TVL1, 400, &p);
    p.lambda = lambda, p.theta = theta, p.tau = tau, p.tol = tol_OF, p.warps = warps;
    float *const xi[4] = {xi11, xi12, xi21, xi22};
    return solve_split(p, nx, ny, I0, I1, nullptr, nullptr, u1, u2, nullptr, xi, verbose, stderr,
                       "Warping: %d,Iter: %d Error: %f\n");
}

extern "C" int faldoi_tvcsad_PD(const float *I0, float *I1, float *xi11, float *xi12, float *xi21, float *xi22, float lambda,
                                float theta, float tau, float tol_OF, int nx, int ny, int warps, int verbose,
                                float *u1, float *u2) {
    faldoi_params p;
    faldoi_default_params(FALDOI_M_TVCSAD, 400, &p);
    p.lambda = lambda, p.theta = theta, p.tau = tau, p.tol = tol_OF, p.warps = warps;
    float *const xi[4] = {xi11, xi12, xi21, xi22};
    return solve_split(p, nx, ny, I0, I1, nullptr, nullptr, u1, u2, nullptr, xi, verbose, stderr,
                       "Warping: %d,Iter: %d Error: %f\n");
}

extern "C" int faldoi_nltvl1_PD(const float *I0, float *I1, float *a, int pd, float lambda, float theta, float tau,
                                int w, int h, int warps, int verbose, float *u1, float *u2) {
    if (pd != 3) {
        set_error("NLTV needs the 3-channel Lab image (pd == 3)");
        return FALDOI_ERR_ARG;
    }
    faldoi_params p;
    faldoi_default_params(FALDOI_M_NLTVL1, 400, &p);
    p.lambda = lambda, p.theta = theta, p.tau = tau, p.warps = warps;
    return solve_split(p, w, h, I0, I1, nullptr, a, u1, u2, nullptr, nullptr, verbose, stdout, "Warping: %d,Iter: %d Error: %f\n");
}

extern "C" int faldoi_nltvcsad_PD(const float *I0, float *I1, float *a, int pd, float lambda, float theta, float tau,
                                  int w, int h, int warps, int verbose, float *u1, float *u2) {
    if (pd != 3) {
        set_error("NLTV needs the 3-channel Lab image (pd == 3)");
        return FALDOI_ERR_ARG;
    }
    faldoi_params p;
    faldoi_default_params(FALDOI_M_NLTVCSAD, 400, &p);
    p.lambda = lambda, p.theta = theta, p.tau = tau, p.warps = warps;
    return solve_split(p, w, h, I0, I1, nullptr, a, u1, u2, nullptr, nullptr, verbose, stdout, "Warping: %d,Iter: %d Error: %f\n");
}

extern "C" int faldoi_guided_tvl2coupled_occ(const float *I0, const float *I1, const float *I_1, float *u1, float *u2,
                                             float *chi, const faldoi_params *p, int nx, int ny, int verbose) {
    if (!p || p->method != FALDOI_M_TVL1_OCC) {
        set_error("faldoi_guided_tvl2coupled_occ: params.method must be 8");
        return FALDOI_ERR_ARG;
    }
    return solve_split(*p, nx, ny, I0, I1, I_1, nullptr, u1, u2, chi, nullptr, verbose, stdout,
                       "Warping: %d, Iter: %d Error: %f\n");
}

// ---------------------------------------------------------------------------
// standalone primitives on host arrays
// ---------------------------------------------------------------------------
namespace {
struct Scratch {
    std::vector<float *> d;
    ~Scratch() {
        for (float *p : d) cudaFree(p);
    }
    float *get(size_t n) {
        float *p = nullptr;
        if (cudaMalloc(&p, n * sizeof(float)) != cudaSuccess) return nullptr;
        d.push_back(p);
        return p;
    }
};
}  // namespace

extern "C" int faldoi_centered_gradient(int device, const float *in, float *dx, float *dy, int nx, int ny) {
    if (!in || !dx || !dy || nx < 2 || ny < 2) return FALDOI_ERR_ARG;
    FALDOI_CUDA(cudaSetDevice(device));
    Geo g = make_geo(nx, ny, nx, 1);
    Scratch sc;
    float *d_in = sc.get(g.plane), *d_dx = sc.get(g.plane), *d_dy = sc.get(g.plane);
    if (!d_in || !d_dx || !d_dy) return FALDOI_ERR_MEM;
    FALDOI_CUDA(cudaMemcpy(d_in, in, g.plane * sizeof(float), cudaMemcpyHostToDevice));
    const dim3 blk(32, 8);
    centered_gradient_kernel<<<grid2d(g, blk, 1), blk>>>(d_in, d_dx, d_dy, g);
    FALDOI_CUDA(cudaGetLastError());
    FALDOI_CUDA(cudaMemcpy(dx, d_dx, g.plane * sizeof(float), cudaMemcpyDeviceToHost));
    FALDOI_CUDA(cudaMemcpy(dy, d_dy, g.plane * sizeof(float), cudaMemcpyDeviceToHost));
    return FALDOI_OK;
}

extern "C" int faldoi_bicubic_warp(int device, const float *in, const float *u, const float *v, float *out, int nx,
                                   int ny, int border_out) {
    if (!in || !u || !v || !out || nx < 1 || ny < 1) return FALDOI_ERR_ARG;
    FALDOI_CUDA(cudaSetDevice(device));
    Geo g = make_geo(nx, ny, nx, 1);
    Scratch sc;
    float *d_in = sc.get(g.plane), *d_u = sc.get(g.plane), *d_v = sc.get(g.plane), *d_o = sc.get(g.plane);
    if (!d_in || !d_u || !d_v || !d_o) return FALDOI_ERR_MEM;
    FALDOI_CUDA(cudaMemcpy(d_in, in, g.plane * sizeof(float), cudaMemcpyHostToDevice));
    FALDOI_CUDA(cudaMemcpy(d_u, u, g.plane * sizeof(float), cudaMemcpyHostToDevice));
    FALDOI_CUDA(cudaMemcpy(d_v, v, g.plane * sizeof(float), cudaMemcpyHostToDevice));
    const dim3 blk(32, 8);
    bicubic_warp_kernel<<<grid2d(g, blk, 1), blk>>>(d_in, d_u, d_v, d_o, 1.f, border_out, g);
    FALDOI_CUDA(cudaGetLastError());
    FALDOI_CUDA(cudaMemcpy(out, d_o, g.plane * sizeof(float), cudaMemcpyDeviceToHost));
    return FALDOI_OK;
}

#include "stripes.inc"

// Self-test of the shared-reciprocal division against IEEE division (see common.cuh).
extern "C" int faldoi_selftest_division(int device, unsigned long long n, unsigned long long seed, unsigned long long *mismatches) {
    if (!mismatches) return FALDOI_ERR_ARG;
    FALDOI_CUDA(cudaSetDevice(device));
    unsigned long long *d = nullptr;
    FALDOI_CUDA(cudaMalloc(&d, sizeof(*d)));
    FALDOI_CUDA(cudaMemset(d, 0, sizeof(*d)));
    selftest_division_kernel<<<148 * 8, 256>>>(n, seed, d);
    cudaError_t e = cudaMemcpy(mismatches, d, sizeof(*d), cudaMemcpyDeviceToHost);
    cudaFree(d);
    FALDOI_CUDA(e);
    return FALDOI_OK;
}

