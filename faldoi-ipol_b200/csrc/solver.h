// Host-side state of one batched solver handle (C++ side of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/faldoi_gpu.h"
#include "common.cuh"
#include "tma.cuh"
#include "tv_tile2_kernel.cuh"
#include "tv_csad2_kernel.cuh"
#include "nltv_tile_kernel.cuh"

namespace faldoi {

void set_error(const std::string &msg);
bool cuda_ok(cudaError_t e, const char *what);

#define FALDOI_CUDA(call)                                   \
    do {                                                    \
        if (!::faldoi::cuda_ok((call), #call)) return FALDOI_ERR_CUDA; \
    } while (0)

enum Family { FAM_TV = 0, FAM_NLTV = 1, FAM_OCC = 2 };

inline bool method_is_csad(int m) { return m == 4 || m == 5 || m == 6 || m == 7; }
inline bool method_is_nltv(int m) { return m == 2 || m == 3 || m == 6 || m == 7; }
inline Family method_family(int m) { return m == 8 ? FAM_OCC : (method_is_nltv(m) ? FAM_NLTV : FAM_TV); }

}  // namespace faldoi

struct faldoi_solver {
    int device = 0, method = 0, B = 0;
    faldoi::Geo g{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<void *> allocs;  // everything cudaMalloc'ed, freed in destroy

    // static per-pair planes
    float *I0 = nullptr, *I1 = nullptr, *I1x = nullptr, *I1y = nullptr;
    size_t i1_plane = 0;  // plane stride of I1/I1x/I1y (== g.plane except in stripe mode, where they are full frames)
    // TV / NLTV state (2 ping-pong sets) and per-warp constants
    float *state = nullptr;
    size_t set_stride = 0;
    float *Ix = nullptr, *Iy = nullptr, *rho_c = nullptr, *scale = nullptr, *I1w = nullptr;
    float *csad_blk = nullptr, *csad_sep = nullptr;  // CSAD two-level sorted residual table (csad_select): CSAD_FLOATS per pixel, CSAD_SEPS separator planes
    faldoi::Tile2Maps maps2{};           // ... of the two-iteration TVL2 kernel
    faldoi::Csad2Maps cmaps{};           // ... of the two-iteration TV-CSAD kernel
    float *csad_t1 = nullptr, *csad_t2 = nullptr;  // TV-CSAD per-warp constants Ix*u1_0, Iy*u2_0
    unsigned *csad_perm = nullptr;                 // TV-CSAD: rank-ordered neighbour codes, 48 bytes per pixel
    double *csad_partial = nullptr;                // TV-CSAD: per-CTA partial error sums [B][2][CTAs per pair]
    unsigned *csad_ticket = nullptr;               // TV-CSAD: [B][t2_stride] arrival counters of the ordered sum
    unsigned char *t2_stat = nullptr;    // [B][t2_stride] launch status of the two-iteration kernel
    int t2_stride = 0;
    // NLTV: Lab, weights, duals
    float *lab = nullptr, *wgt = nullptr, *wt = nullptr, *rwt = nullptr, *dual = nullptr;
    faldoi::NlTileMaps nlmaps{};  // TMA descriptors of nltv_tile_kernel
    size_t dual_set_stride = 0;
    std::vector<int> nl_parity;  // NLTV: ping-pong set that holds each slot's current state (host copy; all pairs run all iterations)
    bool nltv_fast = false;  // faldoi_solver_set_nltv_fast: approximate divisions + paired slot order (not bit-exact)
    // device-side preprocessing (upload_raw): staging for the raw frames, scratch planes, min/max keys
    float *raw_stage = nullptr;
    size_t raw_cap = 0;
    float *pp_tmp = nullptr, *pp_im1 = nullptr;
    unsigned *pp_mm = nullptr;
    // OCC planes
    float *occ = nullptr;  // see occ_kernels.cuh for the layout
    // control
    unsigned *err_max = nullptr;
    double *err_sum = nullptr;
    int *parity = nullptr;
    int *log_iters = nullptr;
    float *log_err = nullptr;
    int err_cap = 0;  // max_iters capacity of err arrays
    // packed export buffer [B][3][h][w] (u1,u2,chi) for plain D2H copies
    float *packed = nullptr;

    // verified constant-divisor (theta) fast division, cached per theta value
    faldoi::DivConst dc{0.f, 0.f, 0};
    bool dc_valid = false;
    int *dc_flag = nullptr;

    // early termination: every CHUNK launches a 1-thread kernel counts the pairs still iterating;
    // the host looks at the count two chunks behind the one it is enqueuing
    int *d_active = nullptr, *h_active = nullptr;  // [4] device / pinned host
    cudaEvent_t chunk_ev[4] = {nullptr, nullptr, nullptr, nullptr};

    float last_ms = 0.f;
    float last_iter_ms = 0.f;             // time inside the per-iteration launches only
    std::vector<cudaEvent_t> phase_ev;    // pairs (begin,end) around each warp's iteration loop
    int phase_used = 0;
    long long launches = 0;
    bool ran = false;

    float *dmalloc(size_t nfloats);
    int phase_mark();  // record the next phase event on the stream
    int alloc_err(int max_iters);
};
