// Shared-memory tiled version of the fused TV iteration (same math and interface as
// tv_iter_kernel in tv_kernels.cuh; see there for the reference citations).
//
// One CTA = one 128 x TT_H pixel tile of one pair (TT_H = 8: 51 KB of shared memory, 256 threads,
// 64 registers -> 4 CTAs = 32 warps per SM; measured on B200: 128x16 tiles at 2 CTAs/SM reach 0.61 of
// the HBM roofline, 128x8 at 4 CTAs/SM 0.71-0.76, see profiles/README.md).  All 11 input planes of
// the tile are staged into shared memory with the bulk async-copy engine (cp.async.bulk ->
// UBLKCP, one 16-byte-aligned row segment per copy, completion on an mbarrier), so no registers are
// held while HBM latency is outstanding and the CTAs of an SM overlap one tile's loads with the
// others' arithmetic:
//   ubar1,ubar2     rows y0-1 .. y0+TT_H,   cols x0-4 .. x0+131   (forward differences + halo)
//   xi11..xi22      rows y0-1 .. y0+TT_H-1, cols x0-4 .. x0+131   (old duals incl. top/left halo)
//   u1,u2,c0,Ix,Iy  rows y0   .. y0+TT_H-1, cols x0   .. x0+127   (c0 = rho_c or the CSAD scale)
// phase 1: xi_new on the tile plus its top row and left column, in place in smem
// phase 2: divergence from smem, data term, primal step, extrapolation, error; the 8 output
//          planes go straight to HBM as float4.
// Out-of-image halo rows / columns are never read by the boundary-aware stencils; the row
// copies that would fall outside the plane are skipped (rows) or land in the allocation's
// guard bands (columns), see faldoi_solver_create.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)

#include "common.cuh"
#include "tv_kernels.cuh"

namespace faldoi {

#ifndef FALDOI_TT_H
#define FALDOI_TT_H 8
#endif
#ifndef FALDOI_TT_CTAS_CSAD
#define FALDOI_TT_CTAS_CSAD 3
#endif
#ifndef FALDOI_TT_CTAS
#define FALDOI_TT_CTAS 4
#endif
enum { TT_W = 128, TT_H = FALDOI_TT_H, TT_PW = TT_W + 8, TT_THREADS = 32 * TT_H };  // one phase-2 quad per thread  // tile size, padded smem row (cols x0-4 .. x0+131)

// Shared-memory planes are TMA box destinations: dense [rows][TT_PW] (or [TT_H][TT_W]), each
// plane padded to a multiple of 128 bytes so every box lands on a 128-byte boundary.
enum {
    TT_UB_ROWS = TT_H + 2,
    TT_XI_ROWS = TT_H + 1,
    TT_UB_FLOATS = (TT_UB_ROWS * TT_PW + 31) / 32 * 32,
    TT_XI_FLOATS = (TT_XI_ROWS * TT_PW + 31) / 32 * 32,
    TT_PL_FLOATS = TT_H * TT_W,
    TT_TX_BYTES = (2 * TT_UB_ROWS * TT_PW + 4 * TT_XI_ROWS * TT_PW + 5 * TT_H * TT_W) * 4
};

struct TileSmem {
    float ub_[2][TT_UB_FLOATS];  // rows y0-1 .. y0+TT_H,   cols x0-4 .. x0+131
    float xi_[4][TT_XI_FLOATS];  // rows y0-1 .. y0+TT_H-1, cols x0-4 .. x0+131
    float pl_[5][TT_PL_FLOATS];  // u1, u2, c0, Ix, Iy: rows y0 .. y0+TT_H-1, cols x0 .. x0+127
    float red[TT_THREADS / 32];
    double redd[TT_THREADS / 32];
    unsigned long long bar;
    __device__ __forceinline__ float *ub(int k, int row) { return &ub_[k][row * TT_PW]; }
    __device__ __forceinline__ float *xi(int k, int row) { return &xi_[k][row * TT_PW]; }
    __device__ __forceinline__ float *pl(int k, int row) { return &pl_[k][row * TT_W]; }
};

// One tensor map per (array, box shape).  All are 3-D {x = pitch, y = rows, z = plane index};
// out-of-range rows / columns of a box are zero-filled by the TMA unit, which is all the frame
// borders need (the boundary-aware stencils never use those values).
struct TileMaps {
    CUtensorMap ub, xi, pl;  // the state array [2 sets][ST_COUNT][B]: boxes 136x(H+2), 136x(H+1), 128xH
    CUtensorMap c0, ix, iy;  // per-warp constants [B]: box 128xH
    CUtensorMap sep;         // CSAD level-1 separators [CSAD_SEPS][B]: box 128xH
};

// TV-CSAD additionally stages the separator planes of the two-level rank table
struct TileSmemCsad : TileSmem {
    alignas(128) float sep_[CSAD_SEPS][TT_PL_FLOATS];
    __device__ __forceinline__ float *sep(int k, int row) { return &sep_[k][row * TT_W]; }
};
template <int DATA>
struct TileSmemFor {
    typedef TileSmem type;
};
// FALDOI_CSAD_SEP_GLOBAL=1 (default): the separators are read with coalesced float4 loads in phase 2;
// 0 stages them with TMA like the other planes (12 KB more shared memory; measured 4 % slower at 16
// pairs).  4 CTAs/SM (64 registers) spills and is 25 % slower, see profiles/README.md.
#ifndef FALDOI_CSAD_SEP_GLOBAL
#define FALDOI_CSAD_SEP_GLOBAL 1
#endif
#if !FALDOI_CSAD_SEP_GLOBAL
template <>
struct TileSmemFor<DATA_CSAD> {
    typedef TileSmemCsad type;
};
#endif

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_box(void *dst_smem, const CUtensorMap *map, int x, int y, int z, unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
            smem_u32(dst_smem)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

template <int DATA>
__global__ void __launch_bounds__(TT_THREADS, DATA == DATA_CSAD ? FALDOI_TT_CTAS_CSAD : FALDOI_TT_CTAS) tv_tile_kernel(const __grid_constant__ TileMaps maps, TvArgs a, int it) {
    // (no pointer arithmetic on the base: it would demote every access from LDS/STS to generic LD/ST)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    typedef typename TileSmemFor<DATA>::type Smem;
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int b = blockIdx.z;
    const int par0 = a.parity[b];  // issued together with the error word read by pair_active: one L2 round trip
    if (!pair_active<DATA>(a, b, it)) return;

    const int w = a.g.w, h = a.g.h, pitch = a.g.pitch;
    const int yo = a.g.y_off, hg = a.g.hg;  // global row = local row + yo (stripe mode), else 0 / h
    const int x0 = blockIdx.x * TT_W, y0 = blockIdx.y * TT_H;
    const int tid = threadIdx.x;
    const int rows = min(TT_H, h - y0);  // interior rows of this tile that exist

    const int par = (par0 + it) & 1;
    const size_t plane = a.g.plane, ks = (size_t)a.g.B * plane;
    float *out = a.state + (size_t)(par ^ 1) * a.set_stride + (size_t)b * plane;

    // ---- stage the tile: one thread arms the mbarrier and issues the 11 TMA box loads ----
    const int ub_lo = (y0 > 0) ? -1 : 0;
    const int xi_lo = ub_lo, xi_hi = rows - 1;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&S.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&S.bar)), "r"((unsigned)(TT_TX_BYTES + (DATA == DATA_CSAD && !FALDOI_CSAD_SEP_GLOBAL ? CSAD_SEPS * TT_PL_FLOATS * 4 : 0))) : "memory");
        const int B = a.g.B, zs = par * ST_COUNT * B + b;
        tma_box(S.ub(0, 0), &maps.ub, x0 - 4, y0 - 1, zs + ST_UB1 * B, &S.bar);
        tma_box(S.ub(1, 0), &maps.ub, x0 - 4, y0 - 1, zs + ST_UB2 * B, &S.bar);
#pragma unroll
        for (int k = 0; k < 4; k++) tma_box(S.xi(k, 0), &maps.xi, x0 - 4, y0 - 1, zs + (ST_XI11 + k) * B, &S.bar);
        tma_box(S.pl(0, 0), &maps.pl, x0, y0, zs + ST_U1 * B, &S.bar);
        tma_box(S.pl(1, 0), &maps.pl, x0, y0, zs + ST_U2 * B, &S.bar);
        tma_box(S.pl(2, 0), &maps.c0, x0, y0, b, &S.bar);
        tma_box(S.pl(3, 0), &maps.ix, x0, y0, b, &S.bar);
        tma_box(S.pl(4, 0), &maps.iy, x0, y0, b, &S.bar);
#if !FALDOI_CSAD_SEP_GLOBAL
        if constexpr (DATA == DATA_CSAD) {
#pragma unroll
            for (int k = 0; k < CSAD_SEPS; k++) tma_box(S.sep(k, 0), &maps.sep, x0, y0, k * B + b, &S.bar);
        }
#endif
    }
    // wait for the bytes (phase 0): one thread polls the mbarrier, the other warps sleep on the CTA
    // barrier instead of burning issue slots in a spin loop (the acquire of the poller is carried
    // to them by bar.sync)
    if (tid == 0) {
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

    const float tau = a.tau, l_t = a.l_t;

    // ---- phase 1: dual step on rows -1..rows-1 (relative), column quads -1..31 (quad -1 = cols x0-4..x0-1) ----
    for (int t = tid; t < (TT_H + 1) * 33; t += TT_THREADS) {
        const int r = t / 33 - 1, q = t % 33 - 1;
        if (r < xi_lo || r > xi_hi) continue;
        const int y = y0 + r, cx = 4 * (q + 1);  // smem column of the quad's first pixel
        const int gx0 = x0 + 4 * q;
        if (gx0 + 3 < 0 || gx0 >= w) continue;
        const bool ylast = (y + yo == hg - 1);
        const float4 B1 = *reinterpret_cast<const float4 *>(S.ub(0, r + 1) + cx);
        const float4 B2 = *reinterpret_cast<const float4 *>(S.ub(1, r + 1) + cx);
        const float b1[5] = {B1.x, B1.y, B1.z, B1.w, S.ub(0, r + 1)[cx + 4]};
        const float b2[5] = {B2.x, B2.y, B2.z, B2.w, S.ub(1, r + 1)[cx + 4]};
        float4 N1 = make_float4(0.f, 0.f, 0.f, 0.f), N2 = N1;
        if (!ylast) {
            N1 = *reinterpret_cast<const float4 *>(S.ub(0, r + 2) + cx);
            N2 = *reinterpret_cast<const float4 *>(S.ub(1, r + 2) + cx);
        }
        const float n1[4] = {N1.x, N1.y, N1.z, N1.w}, n2[4] = {N2.x, N2.y, N2.z, N2.w};
        float4 X11 = *reinterpret_cast<const float4 *>(S.xi(0, r + 1) + cx);
        float4 X12 = *reinterpret_cast<const float4 *>(S.xi(1, r + 1) + cx);
        float4 X21 = *reinterpret_cast<const float4 *>(S.xi(2, r + 1) + cx);
        float4 X22 = *reinterpret_cast<const float4 *>(S.xi(3, r + 1) + cx);
        float x11[4] = {X11.x, X11.y, X11.z, X11.w}, x12[4] = {X12.x, X12.y, X12.z, X12.w};
        float x21[4] = {X21.x, X21.y, X21.z, X21.w}, x22[4] = {X22.x, X22.y, X22.z, X22.w};
        // forward differences of ubar (zero on the last column / row of the FRAME), then
        // xi <- (xi + tau*grad) / max(1, |xi_old|).  The divisor is exactly 1 wherever |xi_old| <= 1
        // (x/1 == x), so the IEEE divisions only run for quads that contain a saturated pixel.
        const bool full = (gx0 + 4 < w);  // all four pixels have a right neighbour inside the frame
        float nr1[4], nr2[4];
        float big = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float u1x = (full || gx0 + k < w - 1) ? b1[k + 1] - b1[k] : 0.f;
            const float u2x = (full || gx0 + k < w - 1) ? b2[k + 1] - b2[k] : 0.f;
            const float u1y = ylast ? 0.f : n1[k] - b1[k];
            const float u2y = ylast ? 0.f : n2[k] - b2[k];
            if (DATA == DATA_TVL1) {
                nr1[k] = sqrtf(x11[k] * x11[k] + x12[k] * x12[k] + x21[k] * x21[k] + x22[k] * x22[k]);
                nr2[k] = nr1[k];
            } else {
                nr1[k] = proj_norm_hypot(x11[k], x12[k]);
                nr2[k] = proj_norm_hypot(x21[k], x22[k]);
            }
            big = fmaxf(big, fmaxf(nr1[k], nr2[k]));
            x11[k] = x11[k] + tau * u1x;
            x12[k] = x12[k] + tau * u1y;
            x21[k] = x21[k] + tau * u2x;
            x22[k] = x22[k] + tau * u2y;
        }
        if (big > 1.f) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (DATA == DATA_TVL1) {
                    if (nr1[k] > 1.f) div4_shared(x11[k], x12[k], x21[k], x22[k], nr1[k]);
                } else {
                    // divide by max(1, |xi_old|): x / 1 == x, so only saturated rows divide
                    if (nr1[k] > 1.f) {
                        x11[k] /= nr1[k];
                        x12[k] /= nr1[k];
                    }
                    if (nr2[k] > 1.f) {
                        x21[k] /= nr2[k];
                        x22[k] /= nr2[k];
                    }
                }
            }
        }
        *reinterpret_cast<float4 *>(S.xi(0, r + 1) + cx) = make_float4(x11[0], x11[1], x11[2], x11[3]);
        *reinterpret_cast<float4 *>(S.xi(1, r + 1) + cx) = make_float4(x12[0], x12[1], x12[2], x12[3]);
        *reinterpret_cast<float4 *>(S.xi(2, r + 1) + cx) = make_float4(x21[0], x21[1], x21[2], x21[3]);
        *reinterpret_cast<float4 *>(S.xi(3, r + 1) + cx) = make_float4(x22[0], x22[1], x22[2], x22[3]);
    }
    __syncthreads();

    // ---- phase 2: divergence, data term, primal step, extrapolation ----
    float emax = 0.f;
    double esum = 0.0;
    for (int t = tid; t < TT_H * 32; t += TT_THREADS) {
        const int r = t >> 5, q = t & 31;
        if (r >= rows) continue;
        const int y = y0 + r, cx = 4 * (q + 1), gx0 = x0 + 4 * q;
        if (gx0 >= pitch || y < a.g.own_lo || y >= a.g.own_hi) continue;  // halo rows belong to the neighbour stripe
        const int gy = y + yo;
        const bool interior = gx0 > 0 && gx0 + 4 < w && gy > 0 && gy < hg - 1;  // no frame border in this quad
        const float4 U1 = *reinterpret_cast<const float4 *>(S.pl(0, r) + 4 * q);
        const float4 U2 = *reinterpret_cast<const float4 *>(S.pl(1, r) + 4 * q);
        const float4 C0 = *reinterpret_cast<const float4 *>(S.pl(2, r) + 4 * q);
        const float4 IX = *reinterpret_cast<const float4 *>(S.pl(3, r) + 4 * q);
        const float4 IY = *reinterpret_cast<const float4 *>(S.pl(4, r) + 4 * q);
        const float u1[4] = {U1.x, U1.y, U1.z, U1.w}, u2[4] = {U2.x, U2.y, U2.z, U2.w};
        const float cc[4] = {C0.x, C0.y, C0.z, C0.w};
        const float ix[4] = {IX.x, IX.y, IX.z, IX.w}, iy[4] = {IY.x, IY.y, IY.z, IY.w};
        // CSAD rank selection (csad_select in tv_kernels.cuh): level 1 from the separator planes picks each
        // pixel's block of ranks, the four blocks are gathered with independent loads (one HBM
        // latency for the quad, no dependent chain), level 2 finishes in registers.
        float med[4] = {0.f, 0.f, 0.f, 0.f};
        if constexpr (DATA == DATA_CSAD) {
            float sv[4];
            int jb[4] = {0, 0, 0, 0}, np[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                sv[k] = (ix[k] * u1[k] + iy[k] * u2[k]) / cc[k];
                np[k] = csad_count(min(gx0 + k, w - 1), gy, w, hg);
            }
#pragma unroll
            for (int j = 0; j < CSAD_SEPS; j++) {
#if FALDOI_CSAD_SEP_GLOBAL
                const float4 E = __ldg(reinterpret_cast<const float4 *>(a.sep + ((size_t)j * a.g.B + b) * plane + (size_t)y * pitch + gx0));
#else
                const float4 E = *reinterpret_cast<const float4 *>(S.sep(j, r) + 4 * q);
#endif
                const float e[4] = {E.x, E.y, E.z, E.w};
#pragma unroll
                for (int k = 0; k < 4; k++) jb[k] += csad_sep_false(e[k], j, np[k], sv[k], l_t, cc[k]);
            }
            CsadBlock blk[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (gx0 + k >= w) jb[k] = 0;  // (pitch padding: separators are not written there)
                blk[k] = csad_gather(a.blk, a.g, b, y, min(gx0 + k, w - 1), jb[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) med[k] = csad_finish(blk[k], jb[k], np[k], sv[k], l_t, cc[k]);
        }
        const float4 M11 = *reinterpret_cast<const float4 *>(S.xi(0, r + 1) + cx);
        const float4 M12 = *reinterpret_cast<const float4 *>(S.xi(1, r + 1) + cx);
        const float4 M21 = *reinterpret_cast<const float4 *>(S.xi(2, r + 1) + cx);
        const float4 M22 = *reinterpret_cast<const float4 *>(S.xi(3, r + 1) + cx);
        const float4 T12 = *reinterpret_cast<const float4 *>(S.xi(1, r) + cx);
        const float4 T22 = *reinterpret_cast<const float4 *>(S.xi(3, r) + cx);
        const float l11 = S.xi(0, r + 1)[cx - 1], l21 = S.xi(2, r + 1)[cx - 1];
        const float m11[4] = {M11.x, M11.y, M11.z, M11.w}, m12[4] = {M12.x, M12.y, M12.z, M12.w};
        const float m21[4] = {M21.x, M21.y, M21.z, M21.w}, m22[4] = {M22.x, M22.y, M22.z, M22.w};
        const float p12[4] = {T12.x, T12.y, T12.z, T12.w}, p22[4] = {T22.x, T22.y, T22.z, T22.w};
        float o1[4], o2[4], ob1[4], ob2[4];
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
            float v1, v2;
            if (DATA == DATA_TVL1) {
                // TH (src/global_faldoi.cpp:693-717) as selects: d = s*(Ix,Iy) with s = +-l_t outside the
                // band, 0 where the gradient vanishes, -rho/grad inside (the quotient is simply unused
                // in the other cases)
                const float grad = ix[k] * ix[k] + iy[k] * iy[k];
                const float rho = cc[k] + (ix[k] * u1[k] + iy[k] * u2[k]);
                const float thr = l_t * grad;
                float sc = -rho / grad;
                sc = grad_is_zero(grad) ? 0.f : sc;
                sc = (rho > thr) ? -l_t : sc;
                sc = (rho < -l_t * grad) ? l_t : sc;
                v1 = u1[k] + sc * ix[k];
                v2 = u2[k] + sc * iy[k];
            } else {
                v1 = u1[k];
                v2 = u2[k];
                if (gx < w) {
                    v1 = csad_apply(u1[k], ix[k], med[k], cc[k]);
                    v2 = csad_apply(u2[k], iy[k], med[k], cc[k]);
                }
            }
            o1[k] = u1[k] - tau * (-d1 + div_const(u1[k] - v1, a.dth));
            o2[k] = u2[k] - tau * (-d2 + div_const(u2[k] - v2, a.dth));
            const float e = (o1[k] - u1[k]) * (o1[k] - u1[k]) + (o2[k] - u2[k]) * (o2[k] - u2[k]);
            if (gx < w) {
                emax = fmaxf(emax, e);
                if (DATA == DATA_CSAD) esum += (double)e;
            }
            ob1[k] = 2 * o1[k] - u1[k];
            ob2[k] = 2 * o2[k] - u2[k];
        }
        const size_t o = (size_t)y * pitch + gx0;
        st4(out + ST_XI11 * ks + o, M11);
        st4(out + ST_XI12 * ks + o, M12);
        st4(out + ST_XI21 * ks + o, M21);
        st4(out + ST_XI22 * ks + o, M22);
        st4(out + ST_U1 * ks + o, make_float4(o1[0], o1[1], o1[2], o1[3]));
        st4(out + ST_U2 * ks + o, make_float4(o2[0], o2[1], o2[2], o2[3]));
        st4(out + ST_UB1 * ks + o, make_float4(ob1[0], ob1[1], ob1[2], ob1[3]));
        st4(out + ST_UB2 * ks + o, make_float4(ob2[0], ob2[1], ob2[2], ob2[3]));
    }

    // ---- convergence measure ----
    const int lane = tid & 31, wid = tid >> 5;
    if (DATA == DATA_TVL1) {
        emax = warp_max(emax);
        if (lane == 0) S.red[wid] = emax;
    } else {
        esum = warp_sum(esum);
        if (lane == 0) S.redd[wid] = esum;
    }
    __syncthreads();
    if (tid == 0) {
        if (DATA == DATA_TVL1) {
            float m = S.red[0];
            for (int i = 1; i < TT_THREADS / 32; i++) m = fmaxf(m, S.red[i]);
            atomicMax(a.err_max + (size_t)b * a.max_iters + it, __float_as_uint(m));
        } else {
            double t = S.redd[0];
            for (int i = 1; i < TT_THREADS / 32; i++) t += S.redd[i];
            atomicAdd(a.err_sum + (size_t)b * a.max_iters + it, t);
        }
    }
}

}  // namespace faldoi
