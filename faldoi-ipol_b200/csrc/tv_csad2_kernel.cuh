// TV-CSAD (methods 4,5; tvcsad_PD, src/global_faldoi.cpp:1449-1637): two primal-dual iterations per pass over
// HBM, with the census data term evaluated from shared-memory tiles instead of a per-pixel table in HBM.
//
// The reference sorts, per pixel and iteration, the 2n+1 values {-(b_j - s)} U {(n-2k) l_t scale} and takes
// element n+1 (:1549-1570); b_j (n <= 48 neighbours of the 7x7 window) is constant within a warp, s moves.
// Round 1 stored the b_j SORTED per pixel (192 B) and gathered a 64-byte block per pixel and iteration, which
// HBM serves as a whole 128-byte line: 228 B/px/iter against 84 algorithmic (profiles/README.md).  Here only the
// sort ORDER is stored -- csad_perm_kernel ranks the b_j once per warp and keeps, per pixel, the 48 neighbour
// codes in rank order, one byte each: 12 dense word planes, 48 B/px, read with perfectly coalesced loads -- and
// the values are recomputed where they are needed from tiles of I0 and I1w staged in shared memory by TMA:
//     b_j = (I0[p] - I0[j] - I1w[p] + I1w[j] + Ix u1_0 + Iy u2_0) / scale           (:1514-1534, same association)
// The wanted element is min(a_(m*), t_(m*-1)) with a_m = -(b_(m) - s) ascending in the rank m, t_m descending and
// m* = number of ranks whose predicate a_m >= t_m is false (tv_kernels.cuh, csad_select): the predicate is
// monotone in m, so m* is found by a 6-step binary search whose probes evaluate b at a rank through the stored
// code -- 6 (b, predicate) evaluations per pixel instead of a gather, no HBM traffic beyond the 48 bytes.
//
// Around that, the structure is tv_tile2_kernel's: one CTA = 120 x C2_H tile, 128 staged columns, one warp per
// staged row, phases 1A / 2A (in shared memory) / 1B / 2B (to HBM); the data term of a row is evaluated by the
// warp that owns the row right before its primal step (lane = pixel for the search, lane = quad for the
// stencils, __syncwarp in between).  HBM traffic per pixel and iteration: state 32 B + constants, tiles and
// codes once per two iterations -- about 95 B against round 1's 228.
//
// Exit test (:1543): mean |du|^2 over the frame > tol^2.  The sum is accumulated in double, deterministically
// (per-CTA partial sums added up in a fixed order by the last CTA to finish) -- the reference's own sum is a racy
// `err_D +=` inside an OpenMP loop; launch bookkeeping (two iterations per launch, fix-up of an exit after the first) is
// tv_tile2_kernel's with the mean in place of the maximum.
#pragma once
#include "tv_tile2_kernel.cuh"

namespace faldoi {

#ifndef FALDOI_C2_H
#define FALDOI_C2_H 9
#endif
#ifndef FALDOI_C2_CTAS
#define FALDOI_C2_CTAS 2
#endif
enum {
    C2_H = FALDOI_C2_H,
    C2_W = 120,
    C2_PW = 128,
    C2_WARPS = C2_H + 3,
    C2_THREADS = 32 * C2_WARPS,
    C2_UB_ROWS = C2_H + 4,
    C2_XI_ROWS = C2_H + 3,
    C2_PL_ROWS = C2_H + 2,
    C2_NPL = 7,                                        // u1, u2, scale, Ix, Iy, t1 = Ix u1_0, t2 = Iy u2_0
    C2_TP = 136,                                       // I0 / I1w tiles: cols x0-8 .. x0+127 (3-pixel window apron, 16-byte aligned origin)
    C2_T_ROWS = C2_PL_ROWS + 6,                        // rows y0-4 .. y0+C2_H+3
    C2_T_FLOATS = (C2_T_ROWS * C2_TP + 31) / 32 * 32,  // TMA destinations stay 128-byte aligned
    C2_WORDS = 12,                                     // rank-ordered neighbour codes: 48 bytes per pixel
    C2_TX_BYTES = (2 * C2_UB_ROWS + 4 * C2_XI_ROWS + C2_NPL * C2_PL_ROWS) * C2_PW * 4 + 2 * C2_T_ROWS * C2_TP * 4
};

struct Csad2Smem {
    float ub_[2][C2_UB_ROWS * C2_PW];       // row index = relative row + 2
    float xi_[4][C2_XI_ROWS * C2_PW];       // row index = relative row + 2
    float pl_[C2_NPL][C2_PL_ROWS * C2_PW];  // row index = relative row + 1
    float med_[C2_PL_ROWS * C2_PW];         // the selected element per staged pixel (row index = relative row + 1)
    float i0_[C2_T_FLOATS], iw_[C2_T_FLOATS];
    double red[2][C2_WARPS];
    unsigned long long bar;
    int last;
    __device__ __forceinline__ float *ub(int k, int r) { return &ub_[k][(r + 2) * C2_PW]; }
    __device__ __forceinline__ float *xi(int k, int r) { return &xi_[k][(r + 2) * C2_PW]; }
    __device__ __forceinline__ float *pl(int k, int r) { return &pl_[k][(r + 1) * C2_PW]; }
    __device__ __forceinline__ float *med(int r) { return &med_[(r + 1) * C2_PW]; }
};

struct Csad2Maps {
    CUtensorMap ub, xi, pl;           // state array, boxes 128 x {H+4, H+3, H+2}
    CUtensorMap sc, ix, iy, t1, t2;   // per-warp constants, box 128 x (H+2)
    CUtensorMap i0, iw;               // I0 and the warped I1, box 136 x (H+8)
};

struct Csad2Args {
    const unsigned *perm;  // [B][plane][C2_WORDS]: the 48 code bytes of a pixel, contiguous, byte r = the code of rank r
    unsigned char *stat;   // [B][stat_stride]: 1 = that launch ran both iterations normally
    int stat_stride;
    // deterministic error sums: every CTA stores its two partial sums, the last CTA of a pair to finish (ticket
    // counter) adds them up in a fixed order -- the value that feeds the exit test does not depend on scheduling
    double *partial;   // [B][2][CTAs per pair]
    unsigned *ticket;  // [B][stat_stride]
};

// neighbour code of window offset (dy, dx), dy, dx in -3..3: one byte, decoded with a shift and a mask
__host__ __device__ __forceinline__ unsigned csad_code(int dy, int dx) { return (unsigned)(((dy + 3) << 4) | (dx + 3)); }

// ---------------------------------------------------------------------------------------------------------------
// Per-warp constants (:1514-1534): scale = hypot(Ix^2+Iy^2, 0.01), t1 = Ix*u1, t2 = Iy*u2 (the flow at the start
// of the warp), and the rank order of the neighbour residuals b_j -- descending, ties by window position, exactly
// the order round 1's sorted table had (csad_constants_kernel) -- as 48 code bytes.
// ---------------------------------------------------------------------------------------------------------------
struct CsadPermArgs {
    const float *I0, *I1w, *Ix, *Iy;  // [B]
    const float *u1, *u2;             // flow planes base (+parity*set_stride)
    const int *parity;
    size_t set_stride;
    float *scale, *t1, *t2;  // [B]
    unsigned *perm;          // [B][plane][C2_WORDS]
    Geo g;
};

__global__ void __launch_bounds__(128) csad_perm_kernel(CsadPermArgs a) {
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
    const float scale = (float)hypot((double)g2, 0.01);
    const float u1 = a.u1[so + p], u2 = a.u2[so + p];
    const float t1 = ix * u1, t2 = iy * u2;
    a.scale[off + p] = scale;
    a.t1[off + p] = t1;
    a.t2[off + p] = t2;
    const float *I0 = a.I0 + off, *I1w = a.I1w + off;
    const float i0p = I0[p], iwp = I1w[p];

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
    // rank sort (descending, ties by slot) -- 48x48 compares, all in registers; the code of the value of rank r
    // becomes byte r of the pixel's 48-byte record
    unsigned char *rec = reinterpret_cast<unsigned char *>(a.perm) + (off + p) * (4 * C2_WORDS);
    s = 0;
#pragma unroll
    for (int k = -3; k <= 3; k++)
#pragma unroll
        for (int l = -3; l <= 3; l++) {
            if (k == 0 && l == 0) continue;
            const int i = s++;
            int rank = 0;
#pragma unroll
            for (int j = 0; j < 48; j++) rank += (bv[j] > bv[i]) || (bv[j] == bv[i] && j < i);
            rec[rank] = (unsigned char)csad_code(k, l);
        }
}

// ---------------------------------------------------------------------------------------------------------------
// The data term: the element of index n+1 of the sorted candidates (see the header), for C2_ILP pixels of a lane at
// once (independent searches interleaved, so that the shared-memory loads and the division chain of one hide behind
// the other's).
//   i0c / iwc : the pixel's position in the I0 / I1w tiles (neighbours at dy*C2_TP + dx)
//   pw        : its 12 code words, in registers.  The binary search never indexes them dynamically: the probe
//               of a step is the last rank of the lower half of the current window of ranks, a fixed byte, and
//               the window is halved with selects (8 + 4 + 2 + 1 word selects and two shifts per search).
// ---------------------------------------------------------------------------------------------------------------
#ifndef FALDOI_C2_ILP
#define FALDOI_C2_ILP 2
#endif
enum { C2_ILP = FALDOI_C2_ILP };

struct C2Pix {  // one search in flight
    unsigned w[8];  // the current window of code words
    const float *i0c, *iwc;
    float i0p, iwp, t1, t2, s, sc, rs, amin;
    int np, pos;
    bool sc_ok;
};

// probe the rank whose code is `code` (rank pos + step - 1) and return whether the search moves to the upper half
__device__ __forceinline__ bool c2_probe(C2Pix &P, unsigned code, int step, float l_t) {
    const bool valid = P.pos + step <= P.np;  // a probe past the end of the list changes nothing
    const int m = P.pos + step - 1;
    const int o = (int)code + (int)(code >> 4) * (C2_TP - 16) - (3 * C2_TP + 3);  // code = (dy+3)*16 + (dx+3)
    const float num = P.i0p - P.i0c[o] - P.iwp + P.iwc[o] + P.t1 + P.t2;
    float bj;
    // numerator / scale with the refined reciprocal of IEEE division's own fast path (common.cuh), formed once per
    // pixel: scale = hypot(grad^2, 0.01) >= 0.01; zero or out-of-range numerators take plain division
    if (P.sc_ok && fastdiv_nz_ok(num))
        bj = div_by_rcp(num, P.sc, P.rs);
    else
        bj = num / P.sc;
    const float am = -(bj - P.s);
    const float tm = csad_t(P.np, m, l_t, P.sc);
    const bool pr = am >= tm;
    P.amin = (valid && pr) ? am : P.amin;  // true predicates come in descending rank order: the last one is the lowest
    const bool up = valid && !pr;
    P.pos += up ? step : 0;
    return up;
}

__device__ __forceinline__ void c2_search(C2Pix (&P)[C2_ILP], const unsigned (&pw)[C2_ILP][C2_WORDS], float l_t, float (&out)[C2_ILP]) {
    bool up[C2_ILP];
    // ranks 0..63 (48 real): probe rank 31 = byte 3 of word 7
#pragma unroll
    for (int p = 0; p < C2_ILP; p++) up[p] = c2_probe(P[p], pw[p][7] >> 24, 32, l_t);
#pragma unroll
    for (int p = 0; p < C2_ILP; p++)
#pragma unroll
        for (int i = 0; i < 8; i++) P[p].w[i] = up[p] ? (i < 4 ? pw[p][8 + i] : 0u) : pw[p][i];
    // window of 32 ranks in w[0..7]: probe rank 15 = byte 3 of w[3]
#pragma unroll
    for (int p = 0; p < C2_ILP; p++) up[p] = c2_probe(P[p], P[p].w[3] >> 24, 16, l_t);
#pragma unroll
    for (int p = 0; p < C2_ILP; p++)
#pragma unroll
        for (int i = 0; i < 4; i++) P[p].w[i] = up[p] ? P[p].w[4 + i] : P[p].w[i];
    // 16 ranks in w[0..3]: probe rank 7 = byte 3 of w[1]
#pragma unroll
    for (int p = 0; p < C2_ILP; p++) up[p] = c2_probe(P[p], P[p].w[1] >> 24, 8, l_t);
#pragma unroll
    for (int p = 0; p < C2_ILP; p++)
#pragma unroll
        for (int i = 0; i < 2; i++) P[p].w[i] = up[p] ? P[p].w[2 + i] : P[p].w[i];
    // 8 ranks in w[0..1]: probe rank 3 = byte 3 of w[0]
#pragma unroll
    for (int p = 0; p < C2_ILP; p++) up[p] = c2_probe(P[p], P[p].w[0] >> 24, 4, l_t);
#pragma unroll
    for (int p = 0; p < C2_ILP; p++) P[p].w[0] = up[p] ? P[p].w[1] : P[p].w[0];
    // 4 ranks in w[0]: probe rank 1 = byte 1
#pragma unroll
    for (int p = 0; p < C2_ILP; p++) up[p] = c2_probe(P[p], (P[p].w[0] >> 8) & 0xffu, 2, l_t);
#pragma unroll
    for (int p = 0; p < C2_ILP; p++) P[p].w[0] = up[p] ? (P[p].w[0] >> 16) : P[p].w[0];
    // 2 ranks: probe rank 0 = byte 0
#pragma unroll
    for (int p = 0; p < C2_ILP; p++) c2_probe(P[p], P[p].w[0] & 0xffu, 1, l_t);
#pragma unroll
    for (int p = 0; p < C2_ILP; p++) {
        const float tprev = (P[p].pos > 0) ? csad_t(P[p].np, P[p].pos - 1, l_t, P[p].sc) : INFINITY;
        out[p] = fminf(P[p].amin, tprev);
    }
}

// the data term of one staged row (lane = pixel): med(r)[c] for every in-frame pixel of columns [c_lo, c_hi)
__device__ __forceinline__ void c2_select_row(Csad2Smem &S, const TvArgs &a, const Csad2Args &c2, int b, int r, int y, int x0, int lane) {
    const int w = a.g.w, pitch = a.g.pitch;
    const uint4 *prow = reinterpret_cast<const uint4 *>(c2.perm + ((size_t)b * a.g.plane + (size_t)y * pitch) * C2_WORDS);
    const int gy = y + a.g.y_off;
#pragma unroll 1
    for (int pass = 0; pass < C2_PW / (32 * C2_ILP); pass++) {
        C2Pix P[C2_ILP];
        unsigned pw[C2_ILP][C2_WORDS];
        int col[C2_ILP];
        bool in[C2_ILP];
#pragma unroll
        for (int p = 0; p < C2_ILP; p++) {
            const int c = 32 * (C2_ILP * pass + p) + lane, gx = x0 - 4 + c;
            col[p] = c;
            in[p] = (gx >= 0 && gx < w);
            const int gxc = min(max(gx, 0), w - 1);  // out-of-frame lanes search a valid pixel and discard the result
            const uint4 q0 = __ldg(prow + (size_t)gxc * 3), q1 = __ldg(prow + (size_t)gxc * 3 + 1), q2 = __ldg(prow + (size_t)gxc * 3 + 2);
            pw[p][0] = q0.x, pw[p][1] = q0.y, pw[p][2] = q0.z, pw[p][3] = q0.w;
            pw[p][4] = q1.x, pw[p][5] = q1.y, pw[p][6] = q1.z, pw[p][7] = q1.w;
            pw[p][8] = q2.x, pw[p][9] = q2.y, pw[p][10] = q2.z, pw[p][11] = q2.w;
            const int cc = gxc - (x0 - 4);  // the clamped pixel's staged column
            const float u1 = S.pl(0, r)[cc], u2 = S.pl(1, r)[cc], ix = S.pl(3, r)[cc], iy = S.pl(4, r)[cc];
            P[p].sc = S.pl(2, r)[cc];
            P[p].t1 = S.pl(5, r)[cc];
            P[p].t2 = S.pl(6, r)[cc];
            const float sn = ix * u1 + iy * u2;  // 0 wherever the warp left the frame (Ix = Iy = 0): 0 / scale = 0, no slow path
            P[p].s = (sn == 0.f) ? sn : sn / P[p].sc;
            const int tc = (r + 4) * C2_TP + cc + 4;
            P[p].i0c = &S.i0_[tc];
            P[p].iwc = &S.iw_[tc];
            P[p].i0p = S.i0_[tc];
            P[p].iwp = S.iw_[tc];
            P[p].np = csad_count(gxc, gy, w, a.g.hg);
            P[p].pos = 0;
            P[p].amin = INFINITY;
            P[p].rs = rcp_refined(P[p].sc);
            P[p].sc_ok = P[p].sc < 1048576.f;
        }
        float med[C2_ILP];
        c2_search(P, pw, a.l_t, med);
#pragma unroll
        for (int p = 0; p < C2_ILP; p++)
            if (in[p]) S.med(r)[col[p]] = med[p];
    }
}

// dual step of one column quad of relative row r, in place: row-wise projection (tvcsad_getD :1428-1446)
__device__ __forceinline__ void c2_dual_quad(Csad2Smem &S, int r, int qi, int gx0, int gy, int w, int hg, float tau) {
    const int cx = 4 * qi;
    const bool ylast = (gy == hg - 1);
    const float4 B1 = *reinterpret_cast<const float4 *>(S.ub(0, r) + cx);
    const float4 B2 = *reinterpret_cast<const float4 *>(S.ub(1, r) + cx);
    const bool has_r = (cx + 4 < C2_PW);
    const float b1[5] = {B1.x, B1.y, B1.z, B1.w, has_r ? S.ub(0, r)[cx + 4] : 0.f};
    const float b2[5] = {B2.x, B2.y, B2.z, B2.w, has_r ? S.ub(1, r)[cx + 4] : 0.f};
    float4 N1 = make_float4(0.f, 0.f, 0.f, 0.f), N2 = N1;
    if (!ylast) {
        N1 = *reinterpret_cast<const float4 *>(S.ub(0, r + 1) + cx);
        N2 = *reinterpret_cast<const float4 *>(S.ub(1, r + 1) + cx);
    }
    const float n1[4] = {N1.x, N1.y, N1.z, N1.w}, n2[4] = {N2.x, N2.y, N2.z, N2.w};
    const float4 X11 = *reinterpret_cast<const float4 *>(S.xi(0, r) + cx);
    const float4 X12 = *reinterpret_cast<const float4 *>(S.xi(1, r) + cx);
    const float4 X21 = *reinterpret_cast<const float4 *>(S.xi(2, r) + cx);
    const float4 X22 = *reinterpret_cast<const float4 *>(S.xi(3, r) + cx);
    float x11[4] = {X11.x, X11.y, X11.z, X11.w}, x12[4] = {X12.x, X12.y, X12.z, X12.w};
    float x21[4] = {X21.x, X21.y, X21.z, X21.w}, x22[4] = {X22.x, X22.y, X22.z, X22.w};
    const bool full = (gx0 + 4 < w);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float u1x = (full || gx0 + k < w - 1) ? b1[k + 1] - b1[k] : 0.f;
        const float u2x = (full || gx0 + k < w - 1) ? b2[k + 1] - b2[k] : 0.f;
        const float u1y = ylast ? 0.f : n1[k] - b1[k];
        const float u2y = ylast ? 0.f : n2[k] - b2[k];
        const float nr1 = proj_norm_hypot(x11[k], x12[k]), nr2 = proj_norm_hypot(x21[k], x22[k]);
        x11[k] = x11[k] + tau * u1x;
        x12[k] = x12[k] + tau * u1y;
        x21[k] = x21[k] + tau * u2x;
        x22[k] = x22[k] + tau * u2y;
        // divide by max(1, |xi_old|): x / 1 == x, so only saturated rows divide (two quotients, one reciprocal)
        if (nr1 > 1.f) div2_shared(x11[k], x12[k], nr1);
        if (nr2 > 1.f) div2_shared(x21[k], x22[k], nr2);
    }
    *reinterpret_cast<float4 *>(S.xi(0, r) + cx) = make_float4(x11[0], x11[1], x11[2], x11[3]);
    *reinterpret_cast<float4 *>(S.xi(1, r) + cx) = make_float4(x12[0], x12[1], x12[2], x12[3]);
    *reinterpret_cast<float4 *>(S.xi(2, r) + cx) = make_float4(x21[0], x21[1], x21[2], x21[3]);
    *reinterpret_cast<float4 *>(S.xi(3, r) + cx) = make_float4(x22[0], x22[1], x22[2], x22[3]);
}

// divergence, data term (med from shared memory), primal step, extrapolation of one quad; returns the sum of |du|^2
// over the quad's in-frame pixels
__device__ __forceinline__ double c2_primal_quad(Csad2Smem &S, const TvArgs &a, int r, int qi, int gx0, int gy, int w, int hg, float (&o1)[4],
                                                 float (&o2)[4], float (&ob1)[4], float (&ob2)[4]) {
    const int cx = 4 * qi;
    const float tau = a.tau;
    const float4 M11 = *reinterpret_cast<const float4 *>(S.xi(0, r) + cx);
    const float4 M12 = *reinterpret_cast<const float4 *>(S.xi(1, r) + cx);
    const float4 M21 = *reinterpret_cast<const float4 *>(S.xi(2, r) + cx);
    const float4 M22 = *reinterpret_cast<const float4 *>(S.xi(3, r) + cx);
    const float4 T12 = *reinterpret_cast<const float4 *>(S.xi(1, r - 1) + cx);
    const float4 T22 = *reinterpret_cast<const float4 *>(S.xi(3, r - 1) + cx);
    const float l11 = cx ? S.xi(0, r)[cx - 1] : 0.f, l21 = cx ? S.xi(2, r)[cx - 1] : 0.f;
    const float4 U1 = *reinterpret_cast<const float4 *>(S.pl(0, r) + cx);
    const float4 U2 = *reinterpret_cast<const float4 *>(S.pl(1, r) + cx);
    const float4 SC = *reinterpret_cast<const float4 *>(S.pl(2, r) + cx);
    const float4 IX = *reinterpret_cast<const float4 *>(S.pl(3, r) + cx);
    const float4 IY = *reinterpret_cast<const float4 *>(S.pl(4, r) + cx);
    const float4 MD = *reinterpret_cast<const float4 *>(S.med(r) + cx);
    const float m11[4] = {M11.x, M11.y, M11.z, M11.w}, m12[4] = {M12.x, M12.y, M12.z, M12.w};
    const float m21[4] = {M21.x, M21.y, M21.z, M21.w}, m22[4] = {M22.x, M22.y, M22.z, M22.w};
    const float p12[4] = {T12.x, T12.y, T12.z, T12.w}, p22[4] = {T22.x, T22.y, T22.z, T22.w};
    const float u1[4] = {U1.x, U1.y, U1.z, U1.w}, u2[4] = {U2.x, U2.y, U2.z, U2.w};
    const float sc[4] = {SC.x, SC.y, SC.z, SC.w}, md[4] = {MD.x, MD.y, MD.z, MD.w};
    const float ix[4] = {IX.x, IX.y, IX.z, IX.w}, iy[4] = {IY.x, IY.y, IY.z, IY.w};
    const bool interior = gx0 > 0 && gx0 + 4 < w && gy > 0 && gy < hg - 1;
    double esum = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int gx = gx0 + k;
        float d1, d2;
        if (interior) {
            d1 = (m11[k] - (k ? m11[k - 1] : l11)) + (m12[k] - p12[k]);
            d2 = (m21[k] - (k ? m21[k - 1] : l21)) + (m22[k] - p22[k]);
        } else {
            d1 = div_bc(m11[k], k ? m11[k - 1] : l11, m12[k], p12[k], gx, gy, w, hg);
            d2 = div_bc(m21[k], k ? m21[k - 1] : l21, m22[k], p22[k], gx, gy, w, hg);
        }
        float v1 = u1[k], v2 = u2[k];
        const bool in = (gx >= 0 && gx < w);
        if (in) csad_apply2(u1[k], u2[k], ix[k], iy[k], md[k], sc[k], v1, v2);
        // u - v is exactly 0 wherever the data term keeps u (a third of the pixels): 0 / theta = 0, and IEEE division
        // would send that lane -- and its warp -- through the out-of-line slow path
        const float x1 = u1[k] - v1, x2 = u2[k] - v2;
        o1[k] = u1[k] - tau * (-d1 + (x1 == 0.f ? x1 : div_const(x1, a.dth)));
        o2[k] = u2[k] - tau * (-d2 + (x2 == 0.f ? x2 : div_const(x2, a.dth)));
        const float e = (o1[k] - u1[k]) * (o1[k] - u1[k]) + (o2[k] - u2[k]) * (o2[k] - u2[k]);
        if (in) esum += (double)e;
        ob1[k] = 2 * o1[k] - u1[k];
        ob2[k] = 2 * o2[k] - u2[k];
    }
    return esum;
}

// the reference's exit test after iteration `it` of pair b: mean |du|^2 > tol^2 (false = stop)
__device__ __forceinline__ bool c2_goes_on(const TvArgs &a, int b, int it) {
    return (float)a.err_sum[(size_t)b * a.max_iters + it] / (float)(a.g.w * a.g.hg) > a.tol2;
}

__global__ void __launch_bounds__(C2_THREADS, FALDOI_C2_CTAS) tv_csad2_kernel(const __grid_constant__ Csad2Maps maps, TvArgs a, Csad2Args c2, int L) {
    extern __shared__ __align__(1024) unsigned char smem_raw_c2[];
    Csad2Smem &S = *reinterpret_cast<Csad2Smem *>(smem_raw_c2);
    const int b = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
    const int par0 = a.parity[b];
    const int it = 2 * L;  // first iteration of this launch

    // ---- what does this launch do for pair b?  (tv_tile2_kernel's bookkeeping) ----
    int mode, par_in = (par0 + L) & 1;
    {
        unsigned char *st = c2.stat + (size_t)b * c2.stat_stride;
        if (L == 0) {
            mode = (it + 1 < a.max_iters) ? T2_MODE_TWO : T2_MODE_ONE;
        } else if (!st[L - 1]) {
            mode = T2_MODE_SKIP;
        } else if (!c2_goes_on(a, b, it - 2)) {
            mode = T2_MODE_FIXUP;  // redo iteration it-2 alone, from the previous launch's input set
            par_in ^= 1;
        } else if (!c2_goes_on(a, b, it - 1) || it >= a.max_iters) {
            mode = T2_MODE_SKIP;
        } else {
            mode = (it + 1 < a.max_iters) ? T2_MODE_TWO : T2_MODE_ONE;
        }
        if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) st[L] = (mode == T2_MODE_TWO);
    }
    if (mode == T2_MODE_SKIP) return;

    const int w = a.g.w, h = a.g.h, pitch = a.g.pitch, hg = a.g.hg, yo = a.g.y_off;
    const int x0 = blockIdx.x * C2_W, y0 = blockIdx.y * C2_H;
    const int rows = min(C2_H, h - y0);
    const size_t plane = a.g.plane, ks = (size_t)a.g.B * plane;
    float *out = a.state + (size_t)(par_in ^ 1) * a.set_stride + (size_t)b * plane;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&S.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&S.bar)), "r"((unsigned)C2_TX_BYTES) : "memory");
        const int B = a.g.B, zs = par_in * ST_COUNT * B + b;
        tma_box(S.ub_[0], &maps.ub, x0 - 4, y0 - 2, zs + ST_UB1 * B, &S.bar);
        tma_box(S.ub_[1], &maps.ub, x0 - 4, y0 - 2, zs + ST_UB2 * B, &S.bar);
#pragma unroll
        for (int k = 0; k < 4; k++) tma_box(S.xi_[k], &maps.xi, x0 - 4, y0 - 2, zs + (ST_XI11 + k) * B, &S.bar);
        tma_box(S.pl_[0], &maps.pl, x0 - 4, y0 - 1, zs + ST_U1 * B, &S.bar);
        tma_box(S.pl_[1], &maps.pl, x0 - 4, y0 - 1, zs + ST_U2 * B, &S.bar);
        tma_box(S.pl_[2], &maps.sc, x0 - 4, y0 - 1, b, &S.bar);
        tma_box(S.pl_[3], &maps.ix, x0 - 4, y0 - 1, b, &S.bar);
        tma_box(S.pl_[4], &maps.iy, x0 - 4, y0 - 1, b, &S.bar);
        tma_box(S.pl_[5], &maps.t1, x0 - 4, y0 - 1, b, &S.bar);
        tma_box(S.pl_[6], &maps.t2, x0 - 4, y0 - 1, b, &S.bar);
        tma_box(S.i0_, &maps.i0, x0 - 8, y0 - 4, b, &S.bar);
        tma_box(S.iw_, &maps.iw, x0 - 8, y0 - 4, b, &S.bar);
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(&S.bar)), "r"(0), "r"(FALDOI_MBAR_SUSPEND_NS)
                : "memory");
        }
    }
    __syncthreads();

    const float tau = a.tau;
    const int qi = lane, gx0 = x0 - 4 + 4 * lane;   // this lane's column quad
    const bool col_ok = (gx0 + 3 >= 0 && gx0 < w);  // some pixel of the quad is inside the frame
    double esumA = 0.0, esumB = 0.0;

    if (mode == T2_MODE_TWO) {
        // ---- 1A: xi_A, warp wi <-> relative row wi-2 (rows -2 .. C2_H) ----
        {
            const int r = wi - 2, y = y0 + r;
            if (y >= 0 && y < h && col_ok) c2_dual_quad(S, r, qi, gx0, y + yo, w, hg, tau);
        }
        __syncthreads();
        // ---- 2A: data term of the row, then u_A, ubar_A in shared memory; warp wi <-> relative row wi-1 (rows -1 .. C2_H) ----
        if (wi < C2_H + 2) {
            const int r = wi - 1, y = y0 + r;
            if (y >= 0 && y < h) {
                c2_select_row(S, a, c2, b, r, y, x0, lane);
                __syncwarp();
                if (col_ok) {
                    float o1[4], o2[4], ob1[4], ob2[4];
                    const double e = c2_primal_quad(S, a, r, qi, gx0, y + yo, w, hg, o1, o2, ob1, ob2);
                    // the error of iteration A counts each pixel once: only the tile's own pixels
                    if (r >= 0 && r < rows && qi >= 1 && qi <= C2_W / 4 && y >= a.g.own_lo && y < a.g.own_hi) esumA = e;
                    const int cx = 4 * qi;
                    *reinterpret_cast<float4 *>(S.pl(0, r) + cx) = make_float4(o1[0], o1[1], o1[2], o1[3]);
                    *reinterpret_cast<float4 *>(S.pl(1, r) + cx) = make_float4(o2[0], o2[1], o2[2], o2[3]);
                    *reinterpret_cast<float4 *>(S.ub(0, r) + cx) = make_float4(ob1[0], ob1[1], ob1[2], ob1[3]);
                    *reinterpret_cast<float4 *>(S.ub(1, r) + cx) = make_float4(ob2[0], ob2[1], ob2[2], ob2[3]);
                }
            }
        }
        __syncthreads();
    }

    // ---- 1B (or the only iteration): xi, warp wi <-> relative row wi-1 (rows -1 .. rows-1), quads 0 .. 30 ----
    if (wi < C2_H + 1) {
        const int r = wi - 1, y = y0 + r;
        if (y >= 0 && r < rows && col_ok && qi <= C2_W / 4) c2_dual_quad(S, r, qi, gx0, y + yo, w, hg, tau);
    }
    __syncthreads();
    // ---- 2B: the tile itself (warp wi <-> row wi, quads 1 .. 30), results to HBM ----
    if (wi < C2_H) {
        const int r = wi, y = y0 + r;
        if (r < rows && y >= a.g.own_lo && y < a.g.own_hi) {
            c2_select_row(S, a, c2, b, r, y, x0, lane);
            __syncwarp();
            if (qi >= 1 && qi <= C2_W / 4 && gx0 < pitch) {
                float o1[4], o2[4], ob1[4], ob2[4];
                esumB = c2_primal_quad(S, a, r, qi, gx0, y + yo, w, hg, o1, o2, ob1, ob2);
                const int cx = 4 * qi;
                const size_t o = (size_t)y * pitch + gx0;
                st4(out + ST_XI11 * ks + o, *reinterpret_cast<const float4 *>(S.xi(0, r) + cx));
                st4(out + ST_XI12 * ks + o, *reinterpret_cast<const float4 *>(S.xi(1, r) + cx));
                st4(out + ST_XI21 * ks + o, *reinterpret_cast<const float4 *>(S.xi(2, r) + cx));
                st4(out + ST_XI22 * ks + o, *reinterpret_cast<const float4 *>(S.xi(3, r) + cx));
                st4(out + ST_U1 * ks + o, make_float4(o1[0], o1[1], o1[2], o1[3]));
                st4(out + ST_U2 * ks + o, make_float4(o2[0], o2[1], o2[2], o2[3]));
                st4(out + ST_UB1 * ks + o, make_float4(ob1[0], ob1[1], ob1[2], ob1[3]));
                st4(out + ST_UB2 * ks + o, make_float4(ob2[0], ob2[1], ob2[2], ob2[3]));
            }
        }
    }

    // ---- convergence measures: err(it) from 2A, err(it+1) from 2B; a fix-up records nothing ----
    if (mode == T2_MODE_FIXUP) return;
    esumA = warp_sum(esumA);
    esumB = warp_sum(esumB);
    if (lane == 0) {
        S.red[0][wi] = esumA;
        S.red[1][wi] = esumB;
    }
    __syncthreads();
    const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    double *pp = c2.partial + (size_t)b * 2 * ncta;
    if (tid == 0) {
        double sA = 0.0, sB = 0.0;
        for (int i = 0; i < C2_WARPS; i++) {
            sA += S.red[0][i];
            sB += S.red[1][i];
        }
        pp[cta] = sA;
        pp[ncta + cta] = sB;
        __threadfence();
        S.last = (atomicAdd(c2.ticket + (size_t)b * c2.stat_stride + L, 1u) == (unsigned)(ncta - 1));
    }
    __syncthreads();
    if (!S.last) return;
    // the last CTA of this pair: ordered sum of all partial sums (fixed assignment of CTAs to threads, fixed
    // shuffle tree, warps added in order)
    __threadfence();
    double tA = 0.0, tB = 0.0;
    for (int i = tid; i < ncta; i += C2_THREADS) {
        tA += __ldcg(pp + i);
        tB += __ldcg(pp + ncta + i);
    }
    tA = warp_sum(tA);
    tB = warp_sum(tB);
    __syncthreads();
    if (lane == 0) {
        S.red[0][wi] = tA;
        S.red[1][wi] = tB;
    }
    __syncthreads();
    if (tid == 0) {
        double sA = 0.0, sB = 0.0;
        for (int i = 0; i < C2_WARPS; i++) {
            sA += S.red[0][i];
            sB += S.red[1][i];
        }
        double *e = a.err_sum + (size_t)b * a.max_iters + it;
        if (mode == T2_MODE_TWO) {
            e[0] = sA;
            e[1] = sB;
        } else {
            e[0] = sB;
        }
    }
}

}  // namespace faldoi
