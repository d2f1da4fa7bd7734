// Tiled, TMA-pipelined version of the fused NLTV iteration (same math, slot order and
// interface as nltv_iter_kernel in nltv_kernels.cuh; see there for the reference citations).
//
// The per-pixel kernel issues ~190 scattered plane loads per pixel and iteration and spends its time
// on the scoreboard (ncu: 10 of 15.7 cycles per issued instruction are long-scoreboard stalls, 3000
// instructions per pixel, a third of them address arithmetic).  Here one CTA owns a 128 x 8 pixel tile
// of one pair (warp = row, lane = 4-pixel quad) and walks the 24 neighbour slots through a small
// ring of shared-memory buffers filled by TMA box loads:
//   per slot s, offset (k,l):   wgt[s], P[s], Q[s]         128 x 8 box at (x0,   y0)    -- own
//                               P[23-s], Q[23-s]           136 x 8 box at (x0-4, y0+k)  -- the neighbour's
//                                                          reciprocal duals; the row shift is in the box
//                                                          origin, the column shift l is applied when
//                                                          reading shared memory (a box must start on a
//                                                          16-byte boundary of global memory: an odd
//                                                          x origin raises "illegal instruction")
//   once per tile:              ubar1, ubar2, 1/wt         (128+8) x (8+4) apron boxes
// Loads of the following slots are in flight while slot s is computed; out-of-frame parts of a box are
// zero-filled by the TMA unit; slots whose neighbour is outside the frame carry a negative weight and
// are clamped to zero weight, which is the reference's neighbour-in-image test.
// New duals go straight to HBM as float4.  Algorithmic traffic is unchanged (532 B / pixel / iteration);
// what changes is who waits for it.
#pragma once
#include "nltv_kernels.cuh"
#include "tma.cuh"  // tma_box, smem_u32

namespace faldoi {

// Stages of the ring x resident CTAs per SM (register budget).  Measured, 4 pairs, exact / fast mode:
// 2x3 (62 KB, 80 registers) 5.03 / 10.6 Gpix*iter/s, 4x2 (102 KB, 107 registers) 4.76 / 10.7, 3x2 4.80, 2x2 4.84.
#ifndef FALDOI_NLT_STAGES
#define FALDOI_NLT_STAGES 3
#endif
#ifndef FALDOI_NLT_CTAS
#define FALDOI_NLT_CTAS 2
#endif
#ifndef FALDOI_NLT_H
#define FALDOI_NLT_H 8  // tile rows
#endif
#ifndef FALDOI_NLT_V
#define FALDOI_NLT_V 2  // pixels per lane (4: one warp per tile row; 2: two warps per row, half the registers per thread)
#endif
enum {
    NLT_W = 128,
    NLT_H = FALDOI_NLT_H,
    NLT_PW = NLT_W + 8,   // apron tile: cols x0-4 .. x0+131
    NLT_AR = NLT_H + 4,   // apron tile: rows y0-2 .. y0+NLT_H+1
    NLT_V = FALDOI_NLT_V,
    NLT_WPR = 4 / NLT_V,              // warps per tile row
    NLT_WARPS = NLT_H * NLT_WPR,
    NLT_THREADS = 32 * NLT_WARPS,
    NLT_NS = FALDOI_NLT_STAGES,
    NLT_TILE = NLT_W * NLT_H,
    NLT_RTILE = NLT_PW * NLT_H,  // reciprocal-dual box: cols x0-4 .. x0+131
    NLT_APRON = NLT_AR * NLT_PW
};
static_assert(NLT_V == 4 || NLT_V == 2, "FALDOI_NLT_V must be 4 or 2");
static_assert((NLT_APRON * 4) % 128 == 0 && (NLT_RTILE * 4) % 128 == 0, "every TMA destination must keep 128-byte alignment");

struct NlTileSmem {
    float own[NLT_NS][3][NLT_TILE];   // wgt[s], P[s], Q[s]
    float rec[NLT_NS][2][NLT_RTILE];  // P[23-s], Q[23-s], rows shifted by k, cols x0-4 ..
    float ub[2][NLT_APRON];
    float rw[NLT_APRON];
    double red[NLT_WARPS];
    unsigned long long full[NLT_NS], cbar;
};

struct NlTileMaps {
    CUtensorMap dual;  // [2 sets][48][B] planes, box 128 x 8
    CUtensorMap rec;   // same array, box 136 x 8
    CUtensorMap wgt;   // [24][B], box 128 x 8
    CUtensorMap ub;    // state [2 sets][ST_COUNT][B], box 136 x 12
    CUtensorMap rwt;   // [B], box 136 x 12: 1/wt (fast mode)
    CUtensorMap wt;    // [B], box 136 x 12: wt (exact mode)
};

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(FALDOI_MBAR_SUSPEND_NS)
            : "memory");
    }
}

// Order in which the slots are visited.  Dual plane s is read twice per iteration: as slot s of its own
// pixel and as the reciprocal of slot 23-s of a neighbour.  Visiting 0, 23, 1, 22, ... puts the two reads
// one step apart, so the second one is an L2 hit instead of a second trip to HBM (the planes of a
// batch are far larger than L2, and in ascending order the reuse distance is up to 23 steps).
// The non-local divergence is then summed in that order, so only the fast arithmetic mode does this;
// the exact mode keeps the reference's ascending order (nltv_kernels.cuh, "Two arithmetic modes").
template <bool PAIRED>
__host__ __device__ constexpr int nl_step_slot(int step) {
    return PAIRED ? ((step & 1) ? NL_SLOTS - 1 - (step >> 1) : (step >> 1)) : step;
}

// aligned vector access of V consecutive floats (float4 / float2)
template <int V>
__device__ __forceinline__ void nl_ld(const float *p, float (&o)[V]) {
    if (V == 4) {
        const float4 q = *reinterpret_cast<const float4 *>(p);
        o[0] = q.x, o[1] = q.y, o[2 % V] = q.z, o[3 % V] = q.w;
    } else {
        const float2 q = *reinterpret_cast<const float2 *>(p);
        o[0] = q.x, o[1] = q.y;
    }
}
template <int V>
__device__ __forceinline__ void nl_st(float *p, const float (&o)[V]) {
    if (V == 4)
        *reinterpret_cast<float4 *>(p) = make_float4(o[0], o[1], o[2 % V], o[3 % V]);
    else
        *reinterpret_cast<float2 *>(p) = make_float2(o[0], o[1]);
}

// out[i] = row[i + l], l in -2..2 known at compile time after unrolling: two aligned vector loads and a
// static pick instead of V scalar loads (which would be V-way bank conflicts at a lane stride of V)
template <int V>
__device__ __forceinline__ void nl_shifted(const float *row, int l, float (&out)[V]) {
    float m[V];
    nl_ld<V>(row, m);
    if (l == 0) {
#pragma unroll
        for (int i = 0; i < V; i++) out[i] = m[i];
    } else if (l < 0) {
        float lo[V];
        nl_ld<V>(row - V, lo);
#pragma unroll
        for (int i = 0; i < V; i++) out[i] = (i + l < 0) ? lo[(V + i + l) % V] : m[(i + l + V) % V];
    } else {
        float hi[V];
        nl_ld<V>(row + V, hi);
#pragma unroll
        for (int i = 0; i < V; i++) out[i] = (i + l >= V) ? hi[(i + l) % V] : m[(i + l) % V];
    }
}

// wt (a sum of up to 24 weights in (0, 1]) inside the divisor range of the reciprocal-based quotients
__device__ __forceinline__ bool nl_wt_ok(float wt) { return wt >= 0.0009765625f && wt < 1048576.f; }  // [2^-10, 2^20)
// weighted flow difference t = w*(ubar_p - ubar_q): zero or 2^-50 <= |t| <= 2^10
__device__ __forceinline__ bool nl_num_ok(float t) {
    const float a = fabsf(t);
    return a == 0.f || (a >= 8.881784197001252e-16f && a <= 1024.f);
}

// the dual update of one pixel and slot with plain IEEE divisions (operands outside the range test of the fast path)
__device__ __noinline__ float4 nl_slot_update_ieee(float t1, float t2, float wp, float wq, float po, float qo, float pr, float qr, float tau) {
    const float g1 = t1 / wp, g2 = t2 / wp;
    const float h1 = -t1 / wq, h2 = -t2 / wq;
    return make_float4((po + tau * g1) / (1 + tau * fabsf(g1)), (qo + tau * g2) / (1 + tau * fabsf(g2)),
                       (pr + tau * h1) / (1 + tau * fabsf(h1)), (qr + tau * h2) / (1 + tau * fabsf(h2)));
}

template <int DATA, bool EXACT>
__global__ void __launch_bounds__(NLT_THREADS, FALDOI_NLT_CTAS) nltv_tile_kernel(const __grid_constant__ NlTileMaps maps, NlArgs a, int it, int base_parity) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    NlTileSmem &S = *reinterpret_cast<NlTileSmem *>(smem_raw);
    constexpr int V = NLT_V;
    const int b = blockIdx.z, tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
    const int r = wi / NLT_WPR, col0 = V * ((wi % NLT_WPR) * 32 + lane);  // tile row, first tile column of this lane's pixels
    const int w = a.g.w, h = a.g.h, pitch = a.g.pitch, B = a.g.B;
    const int x0 = blockIdx.x * NLT_W, y0 = blockIdx.y * NLT_H;
    const int par = (base_parity + it) & 1;
    const int zd = par * 2 * NL_SLOTS * B + b;  // plane index of dual slot 0 of this pair in the input set
    const int zs = par * ST_COUNT * B + b;

    auto issue = [&](int step) {  // one thread: arm the stage's barrier and start its five box loads
        const int sg = step % NLT_NS, s = nl_step_slot<!EXACT>(step), rs = NL_SLOTS - 1 - s;
        int k, l;
        nl_slot_offset(s, k, l);
        (void)l;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&S.full[sg])), "r"((unsigned)((3 * NLT_TILE + 2 * NLT_RTILE) * 4)) : "memory");
        tma_box(S.own[sg][0], &maps.wgt, x0, y0, s * B + b, &S.full[sg]);
        tma_box(S.own[sg][1], &maps.dual, x0, y0, zd + s * B, &S.full[sg]);
        tma_box(S.own[sg][2], &maps.dual, x0, y0, zd + (NL_SLOTS + s) * B, &S.full[sg]);
        tma_box(S.rec[sg][0], &maps.rec, x0 - 4, y0 + k, zd + rs * B, &S.full[sg]);
        tma_box(S.rec[sg][1], &maps.rec, x0 - 4, y0 + k, zd + (NL_SLOTS + rs) * B, &S.full[sg]);
    };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NLT_NS; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&S.full[i])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&S.cbar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&S.cbar)), "r"((unsigned)(3 * NLT_APRON * 4)) : "memory");
        tma_box(S.ub[0], &maps.ub, x0 - 4, y0 - 2, zs + ST_UB1 * B, &S.cbar);
        tma_box(S.ub[1], &maps.ub, x0 - 4, y0 - 2, zs + ST_UB2 * B, &S.cbar);
        tma_box(S.rw, EXACT ? &maps.wt : &maps.rwt, x0 - 4, y0 - 2, b, &S.cbar);
#pragma unroll
        for (int step = 0; step < NLT_NS; step++) issue(step);
    }
    __syncthreads();  // barrier objects are initialised for everyone

    const int y = y0 + r, gx0 = x0 + col0;
    const bool act = (y < h) && (gx0 < w);
    const size_t plane = a.g.plane, ks = (size_t)B * plane, off = (size_t)b * plane;
    const size_t o = (size_t)y * pitch + gx0;
    const float *sin = a.state + (size_t)par * a.set_stride + off;
    float *sout = a.state + (size_t)(par ^ 1) * a.set_stride + off;
    float *dout = a.dual + (size_t)(par ^ 1) * a.dual_set_stride + off;
    const float tau = a.tau, l_t = a.l_t;

    // ---- own pixel data and the data term (plain coalesced float4 loads, once per tile) ----
    float u1[V], u2[V], dv1[V], dv2[V];
#pragma unroll
    for (int i = 0; i < V; i++) u1[i] = u2[i] = dv1[i] = dv2[i] = 0.f;
    if (act) {
        float ix[V], iy[V], cc[V];
        nl_ld<V>(sin + ST_U1 * ks + o, u1);
        nl_ld<V>(sin + ST_U2 * ks + o, u2);
        nl_ld<V>(a.Ix + off + o, ix);
        nl_ld<V>(a.Iy + off + o, iy);
        nl_ld<V>((DATA == DATA_TVL1 ? a.rho_c : a.scale) + off + o, cc);
#pragma unroll
        for (int i = 0; i < V; i++) {
            float v1, v2;
            if (DATA == DATA_TVL1) {
                const float grad = ix[i] * ix[i] + iy[i] * iy[i];
                const float rho = cc[i] + (ix[i] * u1[i] + iy[i] * u2[i]);
                float e1, e2;
                if (rho < -l_t * grad) {
                    e1 = l_t * ix[i];
                    e2 = l_t * iy[i];
                } else if (rho > l_t * grad) {
                    e1 = -l_t * ix[i];
                    e2 = -l_t * iy[i];
                } else if (grad_is_zero(grad)) {
                    e1 = e2 = 0.f;
                } else {
                    const float fi = -rho / grad;
                    e1 = fi * ix[i];
                    e2 = fi * iy[i];
                }
                v1 = u1[i] + e1;
                v2 = u2[i] + e2;
            } else {
                v1 = u1[i];
                v2 = u2[i];
                if (gx0 + i < w && cc[i] != 0.f) {  // 0 marks grad <= GRAD_IS_ZERO (:1734)
                    const float s = (ix[i] * u1[i] + iy[i] * u2[i]) / cc[i];
                    const float med = csad_select(a.blk, a.sep, a.g, b, y, gx0 + i, csad_count(gx0 + i, y, w, h), s, l_t, cc[i]);
                    csad_apply2(u1[i], u2[i], ix[i], iy[i], med, cc[i], v1, v2);
                }
            }
            dv1[i] = div_const(u1[i] - v1, a.dth);
            dv2[i] = div_const(u2[i] - v2, a.dth);
        }
    }

    // ---- apron planes: own ubar and 1/wt ----
    mbar_wait(&S.cbar, 0);
    const int ac = (r + 2) * NLT_PW + 4 + col0;  // this lane's pixels in the apron tiles
    // rwp: 1/wt of the own pixels (fast) or wt itself (exact; 1 in the pitch padding, where wt is 0 and every weight too)
    float c1[V], c2[V], rwp[V], dP[V], dQ[V];
    nl_ld<V>(&S.ub[0][ac], c1);
    nl_ld<V>(&S.ub[1][ac], c2);
    nl_ld<V>(&S.rw[ac], rwp);
    if (EXACT) {
#pragma unroll
        for (int i = 0; i < V; i++) rwp[i] = rwp[i] > 0.f ? rwp[i] : 1.f;
    }
#pragma unroll
    for (int i = 0; i < V; i++) dP[i] = dQ[i] = 0.f;
    // exact mode: wt of the own pixel divides 48 quotients per iteration -- its refined reciprocal (the one IEEE
    // division's own fast path computes) is formed once; okp = wt is in the range where that path is exact
    float rcp_p[V];
    bool okp[V];
#pragma unroll
    for (int i = 0; i < V; i++) rcp_p[i] = 1.f, okp[i] = false;
    if (EXACT) {
#pragma unroll
        for (int i = 0; i < V; i++) {
            okp[i] = nl_wt_ok(rwp[i]);
            rcp_p[i] = rcp_refined(okp[i] ? rwp[i] : 1.f);
        }
    }

    // ---- the 24 slots: dual update + non-local divergence ----
    // The slot loop is rolled over the window rows (5 bodies instead of 24: the fully unrolled exact
    // version is 280 KB of code and stalls on instruction fetch); the column offset l stays a
    // compile-time constant of each body because it selects registers in nl_shifted4.
    auto slot_body = [&](int step, int s, int k, int l) {
        const int sg = step % NLT_NS;
        mbar_wait(&S.full[sg], (step / NLT_NS) & 1);
        if (act) {
            const int tc = r * NLT_W + col0;
            float wv[V], po[V], qo[V], pr[V], qr[V];
            nl_ld<V>(&S.own[sg][0][tc], wv);
            nl_ld<V>(&S.own[sg][1][tc], po);
            nl_ld<V>(&S.own[sg][2][tc], qo);
            nl_shifted<V>(&S.rec[sg][0][r * NLT_PW + 4 + col0], l, pr);
            nl_shifted<V>(&S.rec[sg][1][r * NLT_PW + 4 + col0], l, qr);
            const int nc = ac + k * NLT_PW;  // the lane's columns, neighbour row, in the apron tiles
            float nq1[V], nq2[V], nrw[V];
            nl_shifted<V>(&S.ub[0][nc], l, nq1);
            nl_shifted<V>(&S.ub[1][nc], l, nq2);
            nl_shifted<V>(&S.rw[nc], l, nrw);
            float pn[V], qn[V];
#pragma unroll
            for (int i = 0; i < V; i++) {
                // wgt holds -2 where the neighbour is outside the frame (nltv_init_kernel) and 0 in the pitch
                // padding: clamping at 0 makes such a slot contribute nothing and leave its (zero) dual
                // unchanged, with no per-slot bounds tests.  Everything read for it is finite (zero fill).
                const float wm = fmaxf(wv[i], 0.f);
                const float q1 = nq1[i], q2 = nq2[i];
                const float t1 = wm * (c1[i] - q1), t2 = wm * (c2[i] - q2);
                float Pr, Qr;
                if (EXACT) {
                    // The reference's operations, one for one (ofnltv_getD :1127-1174): g = (w*(u_p-u_q))/wt, then
                    // (sc + tau*g)/(1 + tau*sqrt(g*g)); sqrt(g*g) == |g| whenever g*g is a normal float, and below
                    // that 1 + tau*(either) rounds to 1.  A slot without neighbour divides by 1, not by the zero fill.
                    //
                    // All eight quotients take the instruction sequence of IEEE division's own fast path without its
                    // per-division range check and slow-path call (common.cuh: rcp_refined + div_by_rcp; the two
                    // quotients by the own wt and the two by the neighbour's share a reciprocal).  One range test per
                    // pixel and slot stands in for the eight checks:
                    //   * wt, wt_q in [2^-10, 2^20) and |t| = 0 or in [2^-50, 2^10]  =>  |g|, |h| = 0 or in [2^-70, 2^20]:
                    //     every divisor 1 + tau*|g| is a normal number in [1, 2^21) and no intermediate of the
                    //     first four quotients leaves the normal range;
                    //   * the numerators sc + tau*g: |sc| <= 1 (the update maps [-1,1] into itself), so they are below
                    //     2^21, and a non-zero sum of two floats is at least half an ulp of the smaller-magnitude one,
                    //     i.e. >= 2^-97 here, or it is exactly +0 (round to nearest), which the sequence maps to +0 as
                    //     division does.  A dual never becomes -0 (induction from the +0 start), so zero numerators
                    //     of slots without neighbour, of the padding and of flat regions are exact as well -- and,
                    //     unlike IEEE division, this path has no slow branch for them to drag their warp into.
                    // Anything outside the test takes plain IEEE division.
                    const float wq = wm > 0.f ? nrw[i] : 1.f;
                    const bool fast_ok = okp[i] && nl_wt_ok(wq) && nl_num_ok(t1) && nl_num_ok(t2);
                    if (fast_ok) {
                        const float rq = rcp_refined(wq);
                        const float g1 = div_by_rcp(t1, rwp[i], rcp_p[i]), g2 = div_by_rcp(t2, rwp[i], rcp_p[i]);
                        const float h1 = div_by_rcp(-t1, wq, rq), h2 = div_by_rcp(-t2, wq, rq);
                        const float d1 = 1 + tau * fabsf(g1), d2 = 1 + tau * fabsf(g2), e1 = 1 + tau * fabsf(h1), e2 = 1 + tau * fabsf(h2);
                        pn[i] = div_by_rcp(po[i] + tau * g1, d1, rcp_refined(d1));
                        qn[i] = div_by_rcp(qo[i] + tau * g2, d2, rcp_refined(d2));
                        Pr = div_by_rcp(pr[i] + tau * h1, e1, rcp_refined(e1));
                        Qr = div_by_rcp(qr[i] + tau * h2, e2, rcp_refined(e2));
                    } else {  // out of line: keeps eight IEEE divisions per pixel and slot out of the hot loop's code
                        const float4 r4 = nl_slot_update_ieee(t1, t2, rwp[i], wq, po[i], qo[i], pr[i], qr[i], tau);
                        pn[i] = r4.x, qn[i] = r4.y, Pr = r4.z, Qr = r4.w;
                    }
                } else {
                    const float rwq = nrw[i];
                    // own dual, slot s
                    const float g1 = t1 * rwp[i], g2 = t2 * rwp[i];
                    pn[i] = NL_DIV(po[i] + tau * g1, 1 + tau * fabsf(g1));
                    qn[i] = NL_DIV(qo[i] + tau * g2, 1 + tau * fabsf(g2));
                    // neighbour's reciprocal dual, slot 23-s at q (its difference is the negated one)
                    const float h1 = -t1 * rwq, h2 = -t2 * rwq;
                    Pr = NL_DIV(pr[i] + tau * h1, 1 + tau * fabsf(h1));
                    Qr = NL_DIV(qr[i] + tau * h2, 1 + tau * fabsf(h2));
                }
                dP[i] += wm * (pn[i] - Pr);
                dQ[i] += wm * (qn[i] - Qr);
            }
            nl_st<V>(dout + (size_t)s * ks + o, pn);
            nl_st<V>(dout + (size_t)(NL_SLOTS + s) * ks + o, qn);
        }
        if (step + NLT_NS < NL_SLOTS) {
            __syncthreads();  // everyone is done with this stage's buffers
            if (tid == 0) issue(step + NLT_NS);
        }
    };
    if (EXACT) {  // ascending slots: window rows k = -2..2, columns l = -2..2, centre skipped
        int step = 0;
#pragma unroll 1
        for (int kk = 0; kk < 5; kk++) {
#pragma unroll
            for (int ll = 0; ll < 5; ll++) {
                if (kk == 2 && ll == 2) continue;
                slot_body(step, step, kk - 2, ll - 2);
                step++;
            }
        }
    } else {  // pairs (s, 23-s): offsets (k,l) and (-k,-l)
        int step = 0;
#pragma unroll 1
        for (int kk = 0; kk < 3; kk++) {
#pragma unroll
            for (int ll = 0; ll < 5; ll++) {
                if (kk == 2 && ll >= 2) continue;
                const int sl = kk * 5 + ll;
                slot_body(step, sl, kk - 2, ll - 2);
                slot_body(step + 1, NL_SLOTS - 1 - sl, 2 - kk, 2 - ll);
                step += 2;
            }
        }
    }

    // ---- primal step (+div) and extrapolation ----
    double esum = 0.0;
    if (act) {
        float o1[V], o2[V], b1[V], b2[V];
#pragma unroll
        for (int i = 0; i < V; i++) {
            const float dp = EXACT ? dP[i] / rwp[i] : dP[i] * rwp[i], dq = EXACT ? dQ[i] / rwp[i] : dQ[i] * rwp[i];
            o1[i] = u1[i] - tau * (dp + dv1[i]);
            o2[i] = u2[i] - tau * (dq + dv2[i]);
            if (gx0 + i < w) esum += (double)((o1[i] - u1[i]) * (o1[i] - u1[i]) + (o2[i] - u2[i]) * (o2[i] - u2[i]));
            b1[i] = 2 * o1[i] - u1[i];
            b2[i] = 2 * o2[i] - u2[i];
        }
        nl_st<V>(sout + ST_U1 * ks + o, o1);
        nl_st<V>(sout + ST_U2 * ks + o, o2);
        nl_st<V>(sout + ST_UB1 * ks + o, b1);
        nl_st<V>(sout + ST_UB2 * ks + o, b2);
    }
    // printed error only (the exit test is commented out upstream, :1248)
    esum = warp_sum(esum);
    if (lane == 0) S.red[wi] = esum;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < NLT_WARPS; i++) t += S.red[i];
        atomicAdd(a.err_sum + (size_t)b * a.max_iters + it, t);
    }
}

}  // namespace faldoi
