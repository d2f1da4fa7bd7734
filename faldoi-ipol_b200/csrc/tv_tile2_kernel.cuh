// Temporally blocked TVL2 iteration: TWO primal-dual iterations per pass over HBM.
//
// Same arithmetic as tv_tile_kernel<DATA_TVL1> (see tv_kernels.cuh for the reference citations),
// but a CTA stages its 120x8 tile with a 2-row / 4-column apron, runs iteration A on the apron-extended
// region entirely in shared memory, then iteration B on the tile, and writes the 8 state planes
// once: 41 instead of 76 bytes of HBM traffic per pixel and iteration for 14 % more arithmetic.
//
//   staged (TMA boxes, cols x0-4 .. x0+123):   ubar  rows y0-2 .. y0+9
//                                              xi    rows y0-2 .. y0+8
//                                              u, rho_c, Ix, Iy   rows y0-1 .. y0+8
//   1A  xi_A   rows y0-2 .. y0+8      2A  u_A, ubar_A (in smem)  rows y0-1 .. y0+8, err(2L) over the tile
//   1B  xi_B   rows y0-1 .. y0+7      2B  u_B, ubar_B, xi_B -> HBM, rows y0 .. y0+7, err(2L+1)
//
// Exit test.  Launch L runs iterations 2L and 2L+1 of a warp.  The reference stops after the
// first iteration whose error is <= tol^2; if that is the FIRST iteration of a launch, the launch
// has already applied one iteration too many.  Nothing is lost: the launch read set P and wrote set
// P^1, so set P still holds its input.  The next launch slot sees err(2L) <= tol^2 and, instead of
// iterating, redoes that single iteration from set P into set P^1 ("fix-up", no error update).
// stat[b][L] records whether launch L ran normally, so every launch decides from O(1) words.
#pragma once
#include "tma.cuh"

namespace faldoi {

// Geometry: the staged region is 128 columns (32 float4 quads, x0-4 .. x0+123) so that ONE WARP
// owns one staged row and a lane owns one quad -- no index arithmetic, row tests are warp-uniform,
// and every phase is a single balanced round: H+3 / H+2 / H+1 / H busy warps of the CTA's H+3.  The tile
// written back is the inner 120 x H pixels (quads 1..30).
// FALDOI_T2_H / FALDOI_T2_CTAS: tile rows and resident CTAs per SM.  H = 9: 12 warps, 66 KB, 56 registers ->
// 3 CTAs per SM; a 1024x436 pair is 9 x 49 = 441 CTAs, one wave of the 444 resident slots (H = 8 needs 495:
// a second, nearly empty wave, which is what single-pair latency paid for).  Measured 1 / 2 / 64 pairs:
// H=8 31.3 / 46.2 / 67.7, H=9 39.0 / 48.3 / 70.3, H=10 (48 registers, spills) 33.1 / 40.4 / 61.9 Gpix*iter/s.
#ifndef FALDOI_T2_H
#define FALDOI_T2_H 9
#endif
#ifndef FALDOI_T2_CTAS
#define FALDOI_T2_CTAS 3
#endif
enum {
    T2_H = FALDOI_T2_H,       // tile rows
    T2_W = 120,               // tile columns written by a CTA
    T2_PW = 128,              // staged columns
    T2_QUADS = T2_PW / 4,     // 32
    T2_WARPS = T2_H + 3,      // one per staged xi row
    T2_THREADS = 32 * T2_WARPS,
    T2_UB_ROWS = T2_H + 4,
    T2_XI_ROWS = T2_H + 3,
    T2_PL_ROWS = T2_H + 2,
    T2_UB_FLOATS = T2_UB_ROWS * T2_PW,
    T2_XI_FLOATS = T2_XI_ROWS * T2_PW,
    T2_PL_FLOATS = T2_PL_ROWS * T2_PW,
    T2_TX_BYTES = (2 * T2_UB_ROWS + 4 * T2_XI_ROWS + 5 * T2_PL_ROWS) * T2_PW * 4
};

struct Tile2Smem {
    float ub_[2][T2_UB_FLOATS];  // row index = relative row + 2
    float xi_[4][T2_XI_FLOATS];  // row index = relative row + 2
    float pl_[5][T2_PL_FLOATS];  // u1, u2, rho_c, Ix, Iy; row index = relative row + 1
    float red[2][T2_WARPS];
    unsigned long long bar;
    unsigned item, pad_;  // dataflow grid: the work item this CTA took
    __device__ __forceinline__ float *ub(int k, int r) { return &ub_[k][(r + 2) * T2_PW]; }
    __device__ __forceinline__ float *xi(int k, int r) { return &xi_[k][(r + 2) * T2_PW]; }
    __device__ __forceinline__ float *pl(int k, int r) { return &pl_[k][(r + 1) * T2_PW]; }
};

struct Tile2Maps {
    CUtensorMap ub, xi, pl;  // state array, boxes 128 x {12, 11, 10}
    CUtensorMap c0, ix, iy;  // rho_c, Ix, Iy: box 128 x 10
};

enum { T2_MODE_SKIP = 0, T2_MODE_TWO = 1, T2_MODE_ONE = 2, T2_MODE_FIXUP = 3 };

// dual step of one column quad of relative row r, in place in shared memory
__device__ __forceinline__ void t2_dual_quad(Tile2Smem &S, int r, int qi, int gx0, int gy, int w, int hg, float tau) {
    const int cx = 4 * qi;
    const bool ylast = (gy == hg - 1);
    const float4 B1 = *reinterpret_cast<const float4 *>(S.ub(0, r) + cx);
    const float4 B2 = *reinterpret_cast<const float4 *>(S.ub(1, r) + cx);
    const bool has_r = (cx + 4 < T2_PW);  // the last quad of the staged row has no right neighbour in smem (its value is never needed)
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
    float nr[4];
    float big = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float u1x = (full || gx0 + k < w - 1) ? b1[k + 1] - b1[k] : 0.f;
        const float u2x = (full || gx0 + k < w - 1) ? b2[k + 1] - b2[k] : 0.f;
        const float u1y = ylast ? 0.f : n1[k] - b1[k];
        const float u2y = ylast ? 0.f : n2[k] - b2[k];
        // |xi_old|^2; the square root is only taken where it matters: sqrtf(s) > 1 implies s > 1, and where s > 1 but
        // sqrtf(s) rounds to 1 the division below is by exactly 1
        nr[k] = x11[k] * x11[k] + x12[k] * x12[k] + x21[k] * x21[k] + x22[k] * x22[k];
        big = fmaxf(big, nr[k]);
        x11[k] = x11[k] + tau * u1x;
        x12[k] = x12[k] + tau * u1y;
        x21[k] = x21[k] + tau * u2x;
        x22[k] = x22[k] + tau * u2y;
    }
    if (big > 1.f) {
        // Some pixel of the quad is saturated: divide all four by max(1,|xi|) with the shared-reciprocal
        // fast path (x/1 is exact there too) under ONE range test for the quad; anything unusual
        // (zero / tiny / huge numerator, huge norm) takes IEEE division for the whole quad.
#pragma unroll
        for (int k = 0; k < 4; k++) nr[k] = sqrtf(nr[k]);
        bool ok = big < 1e12f;  // (big holds the largest squared norm)
#pragma unroll
        for (int k = 0; k < 4; k++)
            ok = ok && fastdiv_nz_ok(x11[k]) && fastdiv_nz_ok(x12[k]) && fastdiv_nz_ok(x21[k]) && fastdiv_nz_ok(x22[k]);
        if (ok) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float d = fmaxf(1.f, nr[k]), rr = rcp_refined(d);
                x11[k] = div_by_rcp(x11[k], d, rr);
                x12[k] = div_by_rcp(x12[k], d, rr);
                x21[k] = div_by_rcp(x21[k], d, rr);
                x22[k] = div_by_rcp(x22[k], d, rr);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float d = fmaxf(1.f, nr[k]);
                x11[k] /= d;
                x12[k] /= d;
                x21[k] /= d;
                x22[k] /= d;
            }
        }
    }
    *reinterpret_cast<float4 *>(S.xi(0, r) + cx) = make_float4(x11[0], x11[1], x11[2], x11[3]);
    *reinterpret_cast<float4 *>(S.xi(1, r) + cx) = make_float4(x12[0], x12[1], x12[2], x12[3]);
    *reinterpret_cast<float4 *>(S.xi(2, r) + cx) = make_float4(x21[0], x21[1], x21[2], x21[3]);
    *reinterpret_cast<float4 *>(S.xi(3, r) + cx) = make_float4(x22[0], x22[1], x22[2], x22[3]);
}

// divergence, TH, primal step and extrapolation of one column quad of relative row r.
// Returns max |du|^2 over the quad's pixels that are inside the frame.  Results go to o1,o2,ob1,ob2.
// FAST = true evaluates the three divisions per pixel (TH quotient, two (u-v)/theta) with the
// branch-free fast paths and reports in `unsafe` whether any operand left their validated range;
// the caller then repeats the quad with FAST = false (IEEE division), so the bits never differ.
template <bool FAST>
__device__ __forceinline__ float t2_primal_quad_impl(Tile2Smem &S, const TvArgs &a, int r, int qi, int gx0, int gy, int w, int hg,
                                                     float (&o1)[4], float (&o2)[4], float (&ob1)[4], float (&ob2)[4], bool &unsafe) {
    const int cx = 4 * qi;
    const float tau = a.tau, l_t = a.l_t;
    const float4 M11 = *reinterpret_cast<const float4 *>(S.xi(0, r) + cx);
    const float4 M12 = *reinterpret_cast<const float4 *>(S.xi(1, r) + cx);
    const float4 M21 = *reinterpret_cast<const float4 *>(S.xi(2, r) + cx);
    const float4 M22 = *reinterpret_cast<const float4 *>(S.xi(3, r) + cx);
    const float4 T12 = *reinterpret_cast<const float4 *>(S.xi(1, r - 1) + cx);
    const float4 T22 = *reinterpret_cast<const float4 *>(S.xi(3, r - 1) + cx);
    const float l11 = cx ? S.xi(0, r)[cx - 1] : 0.f, l21 = cx ? S.xi(2, r)[cx - 1] : 0.f;
    const float4 U1 = *reinterpret_cast<const float4 *>(S.pl(0, r) + cx);
    const float4 U2 = *reinterpret_cast<const float4 *>(S.pl(1, r) + cx);
    const float4 C0 = *reinterpret_cast<const float4 *>(S.pl(2, r) + cx);
    const float4 IX = *reinterpret_cast<const float4 *>(S.pl(3, r) + cx);
    const float4 IY = *reinterpret_cast<const float4 *>(S.pl(4, r) + cx);
    const float m11[4] = {M11.x, M11.y, M11.z, M11.w}, m12[4] = {M12.x, M12.y, M12.z, M12.w};
    const float m21[4] = {M21.x, M21.y, M21.z, M21.w}, m22[4] = {M22.x, M22.y, M22.z, M22.w};
    const float p12[4] = {T12.x, T12.y, T12.z, T12.w}, p22[4] = {T22.x, T22.y, T22.z, T22.w};
    const float u1[4] = {U1.x, U1.y, U1.z, U1.w}, u2[4] = {U2.x, U2.y, U2.z, U2.w};
    const float cc[4] = {C0.x, C0.y, C0.z, C0.w};
    const float ix[4] = {IX.x, IX.y, IX.z, IX.w}, iy[4] = {IY.x, IY.y, IY.z, IY.w};
    const bool interior = gx0 > 0 && gx0 + 4 < w && gy > 0 && gy < hg - 1;
    float emax = 0.f;
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
        const float grad = ix[k] * ix[k] + iy[k] * iy[k];
        const float rho = cc[k] + (ix[k] * u1[k] + iy[k] * u2[k]);
        const float thr = l_t * grad;
        const bool gz = grad_is_zero(grad), hi = (rho > thr), lo = (rho < -l_t * grad);
        float sc;
        if (FAST) {
            const bool used = !(gz | hi | lo);       // the quotient is only consumed inside the band
            const float gs = used ? grad : 1.f;      // keeps the unused lanes free of inf / nan
            sc = div_by_rcp(-rho, gs, rcp_refined(gs));
            sc = (rho == 0.f) ? -rho : sc;           // signed zero exactly as IEEE: (-(+-0)) / grad
            unsafe |= used && !(fastdiv_num_ok(rho) && grad < 1e6f);
        } else {
            sc = -rho / grad;
        }
        sc = gz ? 0.f : sc;
        sc = hi ? -l_t : sc;
        sc = lo ? l_t : sc;
        const float v1 = u1[k] + sc * ix[k];
        const float v2 = u2[k] + sc * iy[k];
        const float x1 = u1[k] - v1, x2 = u2[k] - v2;
        float q1, q2;
        if (FAST) {  // Markstein sequence with RN(1/theta), exhaustively verified for this theta (DivConst)
            q1 = __fmaf_rn(__fmaf_rn(-(x1 * a.dth.rb), a.dth.b, x1), a.dth.rb, x1 * a.dth.rb);
            q2 = __fmaf_rn(__fmaf_rn(-(x2 * a.dth.rb), a.dth.b, x2), a.dth.rb, x2 * a.dth.rb);
            q1 = (x1 == 0.f) ? x1 : q1;  // +-0 / theta keeps its sign
            q2 = (x2 == 0.f) ? x2 : q2;
            const float a1 = fabsf(x1), a2 = fabsf(x2);
            unsafe |= (a1 != 0.f && !(a1 > 1e-30f && a1 < 1e30f)) || (a2 != 0.f && !(a2 > 1e-30f && a2 < 1e30f));
        } else {
            q1 = x1 / a.dth.b;
            q2 = x2 / a.dth.b;
        }
        o1[k] = u1[k] - tau * (-d1 + q1);
        o2[k] = u2[k] - tau * (-d2 + q2);
        const float e = (o1[k] - u1[k]) * (o1[k] - u1[k]) + (o2[k] - u2[k]) * (o2[k] - u2[k]);
        if (gx >= 0 && gx < w) emax = fmaxf(emax, e);
        ob1[k] = 2 * o1[k] - u1[k];
        ob2[k] = 2 * o2[k] - u2[k];
    }
    return emax;
}

__device__ __forceinline__ float t2_primal_quad(Tile2Smem &S, const TvArgs &a, int r, int qi, int gx0, int gy, int w, int hg,
                                                float (&o1)[4], float (&o2)[4], float (&ob1)[4], float (&ob2)[4]) {
    bool unsafe = !a.dth.ok;  // theta's reciprocal sequence failed its verification: IEEE division throughout
    float e = 0.f;
    if (!unsafe) e = t2_primal_quad_impl<true>(S, a, r, qi, gx0, gy, w, hg, o1, o2, ob1, ob2, unsafe);
    if (unsafe) e = t2_primal_quad_impl<false>(S, a, r, qi, gx0, gy, w, hg, o1, o2, ob1, ob2, unsafe);
    return e;
}

struct T2Args {
    unsigned char *stat;  // [B][max_iters/2 + 2]: 1 = that launch ran both iterations normally
    int stat_stride;
    // Row-stripe mode (stripes.inc), all zero otherwise:
    int force;          // 0: decide from the error history; T2_MODE_TWO / T2_MODE_ONE: run exactly that, no bookkeeping
    int boundary_first; // CTAs of the tile rows next to a stripe boundary come first in the launch order
    unsigned *sig_cnt;  // [2] arrival counters of the boundary CTAs (towards the upper / lower neighbour)
    unsigned *sig_up, *sig_dn;  // flag words in the neighbour GPUs' memory: "my boundary rows of launch n are in your halo"
    const unsigned *wait_up, *wait_dn;  // the flag words the neighbours write in THIS GPU's memory
    unsigned wait_val;                  // launches every stripe has been asked to complete before this one
};

// Dataflow variant (FLOW): ONE grid runs `flow.iters` forced iterations = several launches' worth of tiles, without a
// grid-wide barrier between them.  A CTA takes a work item from a ticket counter (launch-major, so every item it can
// depend on has been taken by a CTA that is running or done: no deadlock whatever the dispatch order), and waits until
// the 3x3 tiles around its own have completed the launch before - the only tiles whose output it reads and whose input
// it overwrites - which they announce with a release store of their launch sequence number.  The tail of a launch
// (2.2 waves of CTAs on a 270-row stripe) and the launch gap then overlap the next launch instead of idling the GPU.
// Only forced modes (speculative blocks, stripes.inc): the exit test needs the whole frame's maximum.
struct T2Flow {
    unsigned *done;    // [B][nby][nbx]: sequence number + 1 of the last launch the tile has completed (monotone)
    unsigned *ticket;  // work-item counter (monotone, never reset)
    unsigned ticket0;  // its value when this grid starts
    unsigned seq0;     // sequence number of this grid's first launch
    int iters;         // iterations this grid runs: two per launch, the last launch the remainder
    int nbx, nby;
};

template <bool FLOW>
__device__ __forceinline__ void tv_tile2_body(const Tile2Maps &maps, const TvArgs &a, const T2Args &t2, int L, const T2Flow &flow) {
    // (no pointer arithmetic on the base: it would demote every access from LDS/STS to generic LD/ST)
    extern __shared__ __align__(1024) unsigned char smem_raw2[];
    Tile2Smem &S = *reinterpret_cast<Tile2Smem *>(smem_raw2);
    const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
    int b, bx, by, nbx, nby, l = 0;
    if (FLOW) {
        if (tid == 0) S.item = atomicAdd(flow.ticket, 1u) - flow.ticket0;
        __syncthreads();
        nbx = flow.nbx, nby = flow.nby;
        const unsigned ntile = (unsigned)(nbx * nby), per = ntile * (unsigned)a.g.B, item = S.item;
        l = (int)(item / per);
        const unsigned rem = item % per, t = rem % ntile;
        b = (int)(rem / ntile);
        by = (int)(t / (unsigned)nbx), bx = (int)(t % (unsigned)nbx);
        L += l;
    } else {
        b = blockIdx.z, bx = blockIdx.x, by = blockIdx.y, nbx = gridDim.x, nby = gridDim.y;
    }
    const int par0 = a.parity[b];
    const int it = 2 * L;  // first iteration of this launch
    // tile row of this CTA.  Stripe mode: the CTAs that produce a neighbour's halo rows (tile row 0; the last two tile
    // rows) are dispatched first, so that their rows -- and the signal that they are there -- cross NVLink while the
    // interior of the stripe is still being computed.
    if (t2.boundary_first && nby >= 3) by = (by == 0) ? 0 : (by <= 2 ? nby - 3 + by : by - 2);

    // ---- what does this launch do for pair b?  (see the header comment) ----
    int mode, par_in = (par0 + L) & 1;
    if (FLOW) {
        mode = (2 * l + 1 < flow.iters) ? T2_MODE_TWO : T2_MODE_ONE;
    } else if (t2.force) {
        mode = t2.force;
    } else {
        unsigned char *st = t2.stat + (size_t)b * t2.stat_stride;
        if (L == 0) {
            mode = (it + 1 < a.max_iters) ? T2_MODE_TWO : T2_MODE_ONE;
        } else if (!st[L - 1]) {
            mode = T2_MODE_SKIP;
        } else {
            const float e0 = __uint_as_float(a.err_chk[(size_t)b * a.max_iters + it - 2]);
            const float e1 = __uint_as_float(a.err_chk[(size_t)b * a.max_iters + it - 1]);
            if (!(e0 > a.tol2)) {
                mode = T2_MODE_FIXUP;  // redo iteration it-2 alone, from the previous launch's input set
                par_in ^= 1;
            } else if (!(e1 > a.tol2) || it >= a.max_iters) {
                mode = T2_MODE_SKIP;
            } else {
                mode = (it + 1 < a.max_iters) ? T2_MODE_TWO : T2_MODE_ONE;
            }
        }
        if (tid == 0 && bx == 0 && by == 0) st[L] = (mode == T2_MODE_TWO);
    }
    if (mode == T2_MODE_SKIP) return;

    const int w = a.g.w, h = a.g.h, pitch = a.g.pitch, hg = a.g.hg, yo = a.g.y_off;
    const int x0 = bx * T2_W, y0 = by * T2_H;
    const int rows = min(T2_H, h - y0);
    const size_t plane = a.g.plane, ks = (size_t)a.g.B * plane;
    float *out = a.state + (size_t)(par_in ^ 1) * a.set_stride + (size_t)b * plane;

    const unsigned seq = flow.seq0 + (unsigned)l;  // FLOW: launches completed (by everybody) before this one
    if (FLOW && l > 0 && wi == 0) {
        // the launch before is in the same grid: wait for the tiles around this one (lanes 0..8 poll one tile each)
        if (lane < 9) {
            const int ny = by + lane / 3 - 1, nx = bx + lane % 3 - 1;
            if (ny >= 0 && ny < nby && nx >= 0 && nx < nbx) {
                const unsigned *q = flow.done + ((size_t)b * nby + ny) * nbx + nx;
                unsigned v;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(q) : "memory");
                } while ((int)(v - seq) < 0);
            }
        }
        __syncwarp();
        if (lane == 0) __threadfence();
    }
    if (tid == 0) {
        // Stripe mode: a CTA next to a stripe boundary reads halo rows the neighbour GPU stores, and overwrites halo
        // rows the neighbour reads.  Both are safe once the neighbour's boundary CTAs of the launch before have
        // finished, which they announce in a flag word in this GPU's memory: poll it (acquire, system scope) before
        // the tile is loaded.  Interior CTAs -- and the stream -- never wait, so the launches stay back to back.
        if (t2.boundary_first) {
            const unsigned *f = (by == 0) ? t2.wait_up : (by >= nby - 2 ? t2.wait_dn : nullptr);
            const unsigned *f2 = (by == 0 && by >= nby - 2) ? t2.wait_dn : nullptr;  // a stripe of one or two tile rows
            const unsigned want = t2.wait_val + (unsigned)l;
            for (int k = 0; k < 2; k++) {
                const unsigned *q = k ? f2 : f;
                if (!q) continue;
                unsigned v;
                do {
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(q) : "memory");
                } while (v < want);
            }
        }
        // the TMA loads below must see what the flags announced
        if (FLOW || t2.boundary_first) asm volatile("fence.proxy.async;\n" ::: "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&S.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&S.bar)), "r"((unsigned)T2_TX_BYTES) : "memory");
        const int B = a.g.B, zs = par_in * ST_COUNT * B + b;
        tma_box(S.ub_[0], &maps.ub, x0 - 4, y0 - 2, zs + ST_UB1 * B, &S.bar);
        tma_box(S.ub_[1], &maps.ub, x0 - 4, y0 - 2, zs + ST_UB2 * B, &S.bar);
#pragma unroll
        for (int k = 0; k < 4; k++) tma_box(S.xi_[k], &maps.xi, x0 - 4, y0 - 2, zs + (ST_XI11 + k) * B, &S.bar);
        tma_box(S.pl_[0], &maps.pl, x0 - 4, y0 - 1, zs + ST_U1 * B, &S.bar);
        tma_box(S.pl_[1], &maps.pl, x0 - 4, y0 - 1, zs + ST_U2 * B, &S.bar);
        tma_box(S.pl_[2], &maps.c0, x0 - 4, y0 - 1, b, &S.bar);
        tma_box(S.pl_[3], &maps.ix, x0 - 4, y0 - 1, b, &S.bar);
        tma_box(S.pl_[4], &maps.iy, x0 - 4, y0 - 1, b, &S.bar);
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
    const int qi = lane, gx0 = x0 - 4 + 4 * lane;       // this lane's column quad
    const bool col_ok = (gx0 + 3 >= 0 && gx0 < w);      // some pixel of the quad is inside the frame
    float emaxA = 0.f, emaxB = 0.f;

    if (mode == T2_MODE_TWO) {
        // ---- 1A: xi_A, warp wi <-> relative row wi-2 (rows -2 .. T2_H) ----
        {
            const int r = wi - 2, y = y0 + r;
            if (y >= 0 && y < h && col_ok) t2_dual_quad(S, r, qi, gx0, y + yo, w, hg, tau);
        }
        __syncthreads();
        // ---- 2A: u_A, ubar_A in shared memory, warp wi <-> relative row wi-1 (rows -1 .. T2_H) ----
        if (wi < T2_H + 2) {
            const int r = wi - 1, y = y0 + r;
            if (y >= 0 && y < h && col_ok) {
                float o1[4], o2[4], ob1[4], ob2[4];
                const float e = t2_primal_quad(S, a, r, qi, gx0, y + yo, w, hg, o1, o2, ob1, ob2);
                // the error of iteration A counts each pixel once: only the tile's own pixels
                if (r >= 0 && r < rows && qi >= 1 && qi <= T2_W / 4 && y >= a.g.own_lo && y < a.g.own_hi) emaxA = e;
                const int cx = 4 * qi;
                *reinterpret_cast<float4 *>(S.pl(0, r) + cx) = make_float4(o1[0], o1[1], o1[2], o1[3]);
                *reinterpret_cast<float4 *>(S.pl(1, r) + cx) = make_float4(o2[0], o2[1], o2[2], o2[3]);
                *reinterpret_cast<float4 *>(S.ub(0, r) + cx) = make_float4(ob1[0], ob1[1], ob1[2], ob1[3]);
                *reinterpret_cast<float4 *>(S.ub(1, r) + cx) = make_float4(ob2[0], ob2[1], ob2[2], ob2[3]);
            }
        }
        __syncthreads();
    }

    // ---- 1B (or the only iteration): xi, warp wi <-> relative row wi-1 (rows -1 .. rows-1), quads 0 .. 30 ----
    if (wi < T2_H + 1) {
        const int r = wi - 1, y = y0 + r;
        if (y >= 0 && r < rows && col_ok && qi <= T2_W / 4) t2_dual_quad(S, r, qi, gx0, y + yo, w, hg, tau);
    }
    __syncthreads();
    // ---- 2B: the tile itself (warp wi <-> row wi, quads 1 .. 30), results to HBM ----
    if (wi < T2_H) {
        const int r = wi, y = y0 + r;
        if (r < rows && qi >= 1 && qi <= T2_W / 4 && gx0 < pitch && y >= a.g.own_lo && y < a.g.own_hi) {
            float o1[4], o2[4], ob1[4], ob2[4];
            emaxB = t2_primal_quad(S, a, r, qi, gx0, y + yo, w, hg, o1, o2, ob1, ob2);
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
            // Row stripes (one frame over several GPUs): the two owned rows next to a stripe boundary are the
            // neighbour's two halo rows -- store them straight into its planes through the NVLink peer mapping.
            if (a.peer_up && y < a.g.own_lo + 2) {
                float *po = a.peer_up + (size_t)(par_in ^ 1) * a.peer_up_set + (size_t)(a.peer_up_row + y - a.g.own_lo) * pitch + gx0;
                const size_t pp = a.peer_up_plane;
                st4(po + ST_XI11 * pp, *reinterpret_cast<const float4 *>(S.xi(0, r) + cx));
                st4(po + ST_XI12 * pp, *reinterpret_cast<const float4 *>(S.xi(1, r) + cx));
                st4(po + ST_XI21 * pp, *reinterpret_cast<const float4 *>(S.xi(2, r) + cx));
                st4(po + ST_XI22 * pp, *reinterpret_cast<const float4 *>(S.xi(3, r) + cx));
                st4(po + ST_U1 * pp, make_float4(o1[0], o1[1], o1[2], o1[3]));
                st4(po + ST_U2 * pp, make_float4(o2[0], o2[1], o2[2], o2[3]));
                st4(po + ST_UB1 * pp, make_float4(ob1[0], ob1[1], ob1[2], ob1[3]));
                st4(po + ST_UB2 * pp, make_float4(ob2[0], ob2[1], ob2[2], ob2[3]));
            }
            if (a.peer_dn && y >= a.g.own_hi - 2) {
                float *po = a.peer_dn + (size_t)(par_in ^ 1) * a.peer_dn_set + (size_t)(a.peer_dn_row + y - (a.g.own_hi - 2)) * pitch + gx0;
                const size_t pp = a.peer_dn_plane;
                st4(po + ST_XI11 * pp, *reinterpret_cast<const float4 *>(S.xi(0, r) + cx));
                st4(po + ST_XI12 * pp, *reinterpret_cast<const float4 *>(S.xi(1, r) + cx));
                st4(po + ST_XI21 * pp, *reinterpret_cast<const float4 *>(S.xi(2, r) + cx));
                st4(po + ST_XI22 * pp, *reinterpret_cast<const float4 *>(S.xi(3, r) + cx));
                st4(po + ST_U1 * pp, make_float4(o1[0], o1[1], o1[2], o1[3]));
                st4(po + ST_U2 * pp, make_float4(o2[0], o2[1], o2[2], o2[3]));
                st4(po + ST_UB1 * pp, make_float4(ob1[0], ob1[1], ob1[2], ob1[3]));
                st4(po + ST_UB2 * pp, make_float4(ob2[0], ob2[1], ob2[2], ob2[3]));
            }
        }
    }

    // ---- stripe mode: tell the neighbour GPU that this launch's boundary rows are in its halo ----
    // Every boundary CTA has stored its rows through the peer mapping; the last one to arrive (counter in local
    // memory) publishes the launch number with a system-scope release store into the neighbour's flag word, on
    // which the neighbour's next launch waits (stream memory operation) -- no host, no all-to-all.
    if (t2.boundary_first && (by == 0 || by >= nby - 2)) {
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            // FLOW: launches of one grid overlap, so each has its own pair of arrival counters (the last arrival
            // resets them) and the flag is raised with a max: a later launch's signal may overtake an earlier one's
            unsigned *cnt = t2.sig_cnt + (FLOW ? 8 + 2 * l : 0);
            if (by == 0 && t2.sig_up) {
                const unsigned n = atomicAdd(&cnt[0], 1u) + 1u;
                if (FLOW) {
                    if (n == (unsigned)nbx) {
                        cnt[0] = 0;
                        __threadfence_system();
                        asm volatile("red.release.sys.global.max.u32 [%0], %1;\n" ::"l"(t2.sig_up), "r"(t2.wait_val + (unsigned)l + 1u) : "memory");
                    }
                } else if (n % nbx == 0) {
                    __threadfence_system();
                    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(t2.sig_up), "r"(n / nbx) : "memory");
                }
            }
            if (by >= nby - 2 && t2.sig_dn) {
                const unsigned n = atomicAdd(&cnt[1], 1u) + 1u;
                const unsigned per_launch = min(2, nby) * nbx;
                if (FLOW) {
                    if (n == per_launch) {
                        cnt[1] = 0;
                        __threadfence_system();
                        asm volatile("red.release.sys.global.max.u32 [%0], %1;\n" ::"l"(t2.sig_dn), "r"(t2.wait_val + (unsigned)l + 1u) : "memory");
                    }
                } else if (n % per_launch == 0) {
                    __threadfence_system();
                    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(t2.sig_dn), "r"(n / per_launch) : "memory");
                }
            }
        }
    }

    // ---- convergence measures: err(it) from 2A, err(it+1) from 2B; a fix-up records nothing ----
    if (mode == T2_MODE_FIXUP) return;
    emaxA = warp_max(emaxA);
    emaxB = warp_max(emaxB);
    if (lane == 0) {
        S.red[0][wi] = emaxA;
        S.red[1][wi] = emaxB;
    }
    __syncthreads();
    if (tid == 0) {
        float mA = S.red[0][0], mB = S.red[1][0];
        for (int i = 1; i < T2_WARPS; i++) {
            mA = fmaxf(mA, S.red[0][i]);
            mB = fmaxf(mB, S.red[1][i]);
        }
        unsigned *e = a.err_max + (size_t)b * a.max_iters + it;
        if (mode == T2_MODE_TWO) {
            atomicMax(e, __float_as_uint(mA));
            atomicMax(e + 1, __float_as_uint(mB));
        } else {
            atomicMax(e, __float_as_uint(mB));
        }
        if (FLOW) {
            // every thread's stores precede the barrier above: publish "this tile has completed launch seq"
            asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(flow.done + ((size_t)b * nby + by) * nbx + bx), "r"(seq + 1u) : "memory");
        }
    }
}

// one launch = two iterations of every tile, grid (tiles x, tiles y, pairs)
__global__ void __launch_bounds__(T2_THREADS, FALDOI_T2_CTAS) tv_tile2_kernel(const __grid_constant__ Tile2Maps maps, TvArgs a, T2Args t2, int L) {
    tv_tile2_body<false>(maps, a, t2, L, T2Flow{});
}
// the dataflow grid: ceil(flow.iters / 2) launches' worth of tiles, 1-D grid of that many x tiles x pairs CTAs
__global__ void __launch_bounds__(T2_THREADS, FALDOI_T2_CTAS) tv_tile2_flow_kernel(const __grid_constant__ Tile2Maps maps, TvArgs a, T2Args t2, int L,
                                                                                    T2Flow flow) {
    tv_tile2_body<true>(maps, a, t2, L, flow);
}

}  // namespace faldoi
