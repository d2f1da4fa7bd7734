// Fused primal-dual iteration for the TV-regularised models:
//   DATA_TVL1 : tvl2OF     (methods 0,1)  src/global_faldoi.cpp:684-784
//   DATA_CSAD : tvcsad_PD  (methods 4,5)  src/global_faldoi.cpp:1543-1620
//
// One launch = one iteration over all active pairs of the batch.  The
// reference's 7 passes per iteration (TH, 2x forward_gradient, getD,
// 2x divergence, memcpy, getP, extrapolation; 52 words/px) become one pass
// that reads 11 planes and writes 8 (76 B/px): u, ubar, xi x4, and the
// per-warp constants.
//
// Thread mapping: a thread owns 4 consecutive pixels (one float4) of R
// consecutive rows and marches down them; a warp covers 128 columns.  The
// stencil dependencies are resolved in registers:
//   * xi_new(p-1)   (backward x-difference of div): warp shuffle from the
//                    left lane; lane 0 recomputes it from 7 scalar loads,
//   * xi_new(p-w)   (backward y-difference): carried from the previous row of
//                    the march; recomputed once for the row above the strip,
//   * ubar(p+1)      shuffle from the right lane (lane 31: one scalar load),
//   * ubar(p+w)      the next row's float4, reused as the current row next step.
// All state planes are double-buffered (set `par` -> set `par^1`) because
// neighbouring threads need the OLD ubar / xi of pixels another CTA rewrites.
#pragma once
#include "common.cuh"

namespace faldoi {

enum { DATA_TVL1 = 0, DATA_CSAD = 1 };
enum { ST_U1 = 0, ST_U2, ST_UB1, ST_UB2, ST_XI11, ST_XI12, ST_XI21, ST_XI22, ST_COUNT };

struct TvArgs {
    float *state;        // [2 sets][ST_COUNT][B] planes
    size_t set_stride;   // floats between set 0 and set 1
    const float *Ix, *Iy;  // [B] warped gradient of I1
    const float *rho_c;    // [B] TVL1 data term constant
    const float *scale;    // [B] CSAD: hypot(Ix^2+Iy^2, 0.01)
    const float *blk;      // CSAD level 2: per pixel, its neighbour residuals b_j sorted descending, in blocks of
                           // CSAD_BR ranks (see csad_select)
    const float *sep;      // [CSAD_SEPS][B] planes, CSAD level 1: the sorted residual that ends each block
    unsigned *err_max;     // [B][max_iters] float bits of max |du|^2     (DATA_TVL1), written by this handle
    const unsigned *err_chk;  // what the exit test reads: err_max, or the max over all stripes of a stripe group
    double *err_sum;       // [B][max_iters] sum of |du|^2                  (DATA_CSAD)
    const int *parity;     // [B]
    Geo g;
    int max_iters;
    float tau, theta, l_t, tol2;
    DivConst dth;  // division by theta
    // stripe mode (B == 1): neighbours' state arrays, mapped through NVLink peer access.  The rows this
    // handle owns next to a stripe boundary are ALSO stored into the neighbour's halo row by the
    // iteration kernel itself (fused compute + halo exchange, no pack/copy/unpack launches).
    float *peer_up, *peer_dn;          // neighbour's `state` base, or nullptr
    size_t peer_up_plane, peer_dn_plane, peer_up_set, peer_dn_set;
    int peer_up_row, peer_dn_row;      // neighbour-local row index of its halo row
};

// Has pair b already met the reference's exit test `err > tol^2` (false = stop)
// after iteration it-1?  Entries of iterations that never ran stay 0 -> stop.
template <int DATA>
__device__ __forceinline__ bool pair_active(const TvArgs &a, int b, int it) {
    if (it == 0) return true;
    float e;
    if (DATA == DATA_TVL1)
        e = __uint_as_float(a.err_chk[(size_t)b * a.max_iters + it - 1]);
    else
        e = (float)a.err_sum[(size_t)b * a.max_iters + it - 1] / (float)(a.g.w * a.g.h);
    return e > a.tol2;
}

// rank selection of the CSAD data term (src/global_faldoi.cpp:1549-1570): the
// element of index n+1 of sort({-(b_j - s)} U {(n-2k)*l_t*scale, k=0..n}).
// With a_m = -(bs_m - s) ascending (bs sorted descending) and the thresholds
// t_m descending, that element equals min_m max(a_m, t_m) = min(a_{m*}, t_{m*-1}) where
// m* = number of ranks whose predicate a_m >= t_m is false (the predicate is monotone in m;
// DESIGN.md, "CSAD rank").  s moves a lot from one iteration to the next (measured: m* jumps by more
// than 8 ranks for a quarter of the pixels), so instead of a data-dependent chain of probes the
// sorted residuals are kept as a two-level table:
//   level 1  separator planes (the last rank of every block but the final one), dense, streamed
//            with the other planes;
//   level 2  blocks of CSAD_BR consecutive ranks, one aligned slot per block.
// One gather per pixel, no dependent loads, 15 comparisons.  Slots of ranks >= np hold -inf: their
// predicate is true and their a is +inf, which is what the reference's shorter list amounts to.
// Default: 4 blocks of 12 ranks in 64-byte slots, the 4 slots of a pixel adjacent (256 B / pixel).
// FALDOI_CSAD_BR=8 selects 6 blocks of 8 ranks laid out [row][block][x][8], where neighbours that
// pick the same block share a 128-byte line: 21 % less HBM traffic (HBM delivers whole lines either
// way) but measured 20 % slower on B200 (profiles/README.md), so it is not the default.
#ifndef FALDOI_CSAD_BR
#define FALDOI_CSAD_BR 12
#endif
enum {
    CSAD_BR = FALDOI_CSAD_BR,            // ranks per block
    CSAD_BLOCKS = 48 / CSAD_BR,
    CSAD_SEPS = CSAD_BLOCKS - 1,
    CSAD_QV = CSAD_BR / 4,               // float4 per block
    CSAD_FLOATS = CSAD_BR == 12 ? 64 : 48  // table floats per pixel
};
static_assert(CSAD_BR == 12 || CSAD_BR == 8, "FALDOI_CSAD_BR must be 12 or 8");

struct CsadBlock {
    float4 q[CSAD_QV];
};

__device__ __forceinline__ float csad_t(int np, int m, float l_t, float scale) { return (float)(np - 2 * m) * l_t * scale; }

// level 1: the block that holds m* = number of separators whose predicate is false
__device__ __forceinline__ int csad_sep_false(float sep_k, int k, int np, float s, float l_t, float scale) {
    return (-(sep_k - s) >= csad_t(np, CSAD_BR * k + CSAD_BR - 1, l_t, scale)) ? 0 : 1;
}

// level 2: the residuals of block j (ranks CSAD_BR*j ..) -> the selected value
__device__ __forceinline__ float csad_finish(const CsadBlock &B, int j, int np, float s, float l_t, float scale) {
    float v[CSAD_BR];
#pragma unroll
    for (int i = 0; i < CSAD_QV; i++) v[4 * i] = B.q[i].x, v[4 * i + 1] = B.q[i].y, v[4 * i + 2] = B.q[i].z, v[4 * i + 3] = B.q[i].w;
    const float fb = (float)(np - 2 * CSAD_BR * j);  // np - 2*(BR*j+i) = fb - 2i, exact in fp32
    int cnt = 0;
    float amin = INFINITY;  // a_{m*}: the smallest a whose predicate holds (+inf if m* is past the last residual)
#pragma unroll
    for (int i = 0; i < CSAD_BR; i++) {
        const float am = -(v[i] - s);
        const float tm = (fb - (float)(2 * i)) * l_t * scale;
        const bool pr = am >= tm;
        cnt += pr ? 0 : 1;
        amin = pr ? fminf(amin, am) : amin;
    }
    const int m = CSAD_BR * j + cnt;
    const float tprev = (m > 0) ? csad_t(np, m - 1, l_t, scale) : INFINITY;
    return fminf(amin, tprev);
}

// float index of rank 0 of block j of pixel (x, y) of pair b.  plane >= h * pitch.
__device__ __forceinline__ size_t csad_slot_index(const Geo &g, int b, int y, int x, int j) {
    if (CSAD_BR == 12) return ((size_t)b * g.plane + (size_t)y * g.pitch + x) * 64 + j * 16;
    return (size_t)b * g.plane * 48 + ((size_t)y * CSAD_BLOCKS + j) * g.pitch * 8 + (size_t)x * 8;
}

__device__ __forceinline__ CsadBlock csad_gather(const float *__restrict__ blk, const Geo &g, int b, int y, int x, int j) {
    const float4 *q = reinterpret_cast<const float4 *>(blk + csad_slot_index(g, b, y, x, j));
    CsadBlock B;
#pragma unroll
    for (int i = 0; i < CSAD_QV; i++) B.q[i] = __ldg(q + i);
    return B;
}

// whole lookup with plain loads (march kernel, NLTV-CSAD)
__device__ __forceinline__ float csad_select(const float *__restrict__ blk, const float *__restrict__ sep, const Geo &g, int b, int y, int x,
                                             int np, float s, float l_t, float scale) {
    const size_t pix = (size_t)b * g.plane + (size_t)y * g.pitch + x, ss = (size_t)g.B * g.plane;
    int j = 0;
#pragma unroll
    for (int k = 0; k < CSAD_SEPS; k++) j += csad_sep_false(__ldg(sep + k * ss + pix), k, np, s, l_t, scale);
    return csad_finish(csad_gather(blk, g, b, y, x, j), j, np, s, l_t, scale);
}

// v = u - (Ix, Iy) * med / scale of the CSAD data term (:1570, :1757).  med is exactly 0 whenever the rank lands on
// the middle threshold (t_{n/2} = 0: "keep u"), which is common -- a third of the pixels -- and a zero numerator
// sends IEEE division, and with it the whole warp, through its out-of-line slow path (the compiler turns an
// `if (n == 0)` around the division into a select, so the division still executes).  The two quotients share the
// refined reciprocal of scale instead (common.cuh: rcp_refined / div_by_rcp), which has no such path: 0 / scale
// is the signed zero itself, everything else is bit for bit the IEEE quotient; out-of-range operands take `/`.
__device__ __forceinline__ void csad_apply2(float u1, float u2, float ix, float iy, float med, float scale, float &v1, float &v2) {
    const float n1 = ix * med, n2 = iy * med;
    float q1, q2;
    if (scale >= 7.450580596923828e-09f && scale < 1048576.f && fastdiv_num_ok(n1) && fastdiv_num_ok(n2)) {
        const float rs = rcp_refined(scale);
        q1 = div_by_rcp(n1, scale, rs);
        q2 = div_by_rcp(n2, scale, rs);
    } else {
        q1 = n1 / scale;
        q2 = n2 / scale;
    }
    v1 = u1 - (n1 == 0.f ? n1 : q1);
    v2 = u2 - (n2 == 0.f ? n2 : q2);
}

// Norm used by TV-CSAD's row-wise projection, max(1, hypotf(a,b)) (tvcsad_getD :1433-1443).
// Only values > 1 matter; a^2+b^2 evaluated in fp32 is within 1.5 ulp (1.8e-7 relative) of the exact sum, so
// below 1 - 5e-7 the exact sum is below 1 - 3e-7, its square root -- however glibc rounds it -- is below 1,
// and the double-precision path is skipped.
__device__ __forceinline__ float proj_norm_hypot(float a, float b) {
    const float s = a * a + b * b;
    if (s < 0.9999995f) return 1.f;
    return hypotf_exact(a, b);
}

__device__ __forceinline__ int csad_count(int x, int y, int w, int h) {
    const int nx = min(x, 3) + min(w - 1 - x, 3) + 1;
    const int ny = min(y, 3) + min(h - 1 - y, 3) + 1;
    return nx * ny - 1;
}

// ---------------------------------------------------------------------------
// CSAD per-warp constants (src/global_faldoi.cpp:1514-1534): scale =
// hypot(Ix^2+Iy^2, 0.01) and, for the in-image neighbours j of the 7x7 window,
//   b_j = (I0[p] - I0[j] - I1w[p] + I1w[j] + Ix*u1 + Iy*u2) / scale
// stored SORTED DESCENDING per pixel as the two-level table of csad_select (blocks of CSAD_BR
// ranks + separator planes) so the per-iteration rank selection is 15 comparisons
// instead of the reference's std::sort of 97 floats.
// hyp = 0 is the NLTV-CSAD variant (:1698-1723): scale = sqrt(Ix^2+Iy^2),
// only where Ix^2+Iy^2 > 1e-8 (elsewhere scale := 0 marks "v = u").
// ---------------------------------------------------------------------------
struct CsadArgs {
    const float *I0, *I1w, *Ix, *Iy;  // [B]
    const float *u1, *u2;             // flow planes base (+parity*set_stride)
    const int *parity;
    size_t set_stride;
    float *scale;  // [B]
    float *blk;    // csad_slot_index layout
    float *sep;    // [CSAD_SEPS][B][plane]
    Geo g;
    int hyp;
};

__global__ void __launch_bounds__(128) csad_constants_kernel(CsadArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    const int w = a.g.w, h = a.g.h, pitch = a.g.pitch;
    if (x >= w || y >= h) return;
    const size_t off = (size_t)b * a.g.plane;
    const size_t so = off + (size_t)a.parity[b] * a.set_stride;
    const int p = y * pitch + x;
    const float ix = a.Ix[off + p], iy = a.Iy[off + p];
    const float g2 = ix * ix + iy * iy;
    float scale;
    if (a.hyp) {
        scale = (float)hypot((double)g2, 0.01);
    } else {
        if (!grad_above_zero(g2)) {
            a.scale[off + p] = 0.f;
            return;
        }
        scale = sqrtf(g2);
    }
    a.scale[off + p] = scale;
    const float u1 = a.u1[so + p], u2 = a.u2[so + p];
    const float *I0 = a.I0 + off, *I1w = a.I1w + off;
    const float i0p = I0[p], iwp = I1w[p];
    const float t1 = ix * u1, t2 = iy * u2;

    float bv[48];
    int s = 0;
#pragma unroll
    for (int k = -3; k <= 3; k++)
#pragma unroll
        for (int l = -3; l <= 3; l++) {
            if (k == 0 && l == 0) continue;
            const int r = y + k, c = x + l;
            float v = -INFINITY;  // out-of-image slots sort to the end
            if (c >= 0 && c < w && r >= 0 && r < h) {
                const int q = r * pitch + c;
                v = (i0p - __ldg(I0 + q) - iwp + __ldg(I1w + q) + t1 + t2) / scale;
            }
            bv[s++] = v;
        }
    // rank sort (descending, ties by slot) -- 48x48 compares, all in registers
#pragma unroll
    for (int i = 0; i < 48; i++) {
        int rank = 0;
#pragma unroll
        for (int j = 0; j < 48; j++) rank += (bv[j] > bv[i]) || (bv[j] == bv[i] && j < i);
        a.blk[csad_slot_index(a.g, b, y, x, rank / CSAD_BR) + rank % CSAD_BR] = bv[i];
        if (rank % CSAD_BR == CSAD_BR - 1 && rank < 47) a.sep[(size_t)(rank / CSAD_BR) * a.g.B * a.g.plane + off + p] = bv[i];
    }
}

}  // namespace faldoi
