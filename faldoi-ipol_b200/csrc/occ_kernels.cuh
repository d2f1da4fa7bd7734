// TVL2 + occlusion model (method 8): guided_tvl2coupled_occ
// src/tvl2_model_occ.cpp:492-779 with the patch = whole image, GLOBAL_STEP.
//
// One outer iteration = v-update, 24 Chambolle sweeps for xi (:340-392) + final
// divergence and u-update (:394-406, :726-751), 24 primal-dual sweeps for chi
// (:431-474) and the hard threshold (:476-483).  The reference materialises
// g*xi, div(g*xi), vi, grad(vi), g*eta, div(g*eta), chix, chiy as planes in
// every sweep; here each sweep is ONE stencil kernel that recomputes those
// intermediates from the radius-1 neighbourhood, so a xi sweep moves 9 planes
// in / 4 out and a chi sweep 6 in / 3 out.
//
// Target definition: eta1 = eta2 = 0 and div_u = 0 at entry -- the
// reference reads them uninitialised on this path.
#pragma once
#include "common.cuh"
#include "warp_kernels.cuh"

namespace faldoi {

enum {
    OC_IM1 = 0, OC_JX, OC_JY, OC_G, OC_UBA1, OC_UBA2, OC_I1WX, OC_I1WY, OC_JWX, OC_JWY, OC_RHO_C1, OC_RHO_C_1,
    OC_U1, OC_U2, OC_V1, OC_V2, OC_K1, OC_K2, OC_F, OC_GG,
    OC_XI0,            // 4 planes, set 0
    OC_XI1 = OC_XI0 + 4,   // 4 planes, set 1
    OC_ETA0 = OC_XI1 + 4,  // 2 planes, set 0
    OC_ETA1 = OC_ETA0 + 2,
    OC_CHI0 = OC_ETA1 + 2,
    OC_CHI1,
    OCC_PLANES
};

struct OccArgs {
    float *pl;            // [OCC_PLANES][B]
    const float *I0, *I1, *I1x, *I1y;
    unsigned *err_max;    // [B][max_iters]
    Geo g;
    int max_iters;
    float lambda, theta, beta, alpha, tau_theta, mu, tau_eta, tau_chi, l_t, tol2;
};

__device__ __forceinline__ float *occ_plane(const OccArgs &a, int kind, int b) {
    return a.pl + ((size_t)kind * a.g.B + b) * a.g.plane;
}

__device__ __forceinline__ bool occ_active(const OccArgs &a, int b, int it) {
    if (it == 0) return true;
    return __uint_as_float(a.err_max[(size_t)b * a.max_iters + it - 1]) > a.tol2;
}

// once per run: gradients of I-1, weight g = 1/(1+0.05*|grad I0|) (init_weight
// src/utils.cpp:838-852), frozen backward flow u_ba = -u (:597-598), xi = 0
__global__ void __launch_bounds__(256) occ_init_kernel(OccArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    const int w = a.g.w, h = a.g.h, pitch = a.g.pitch;
    if (x >= w || y >= h) return;
    const int p = y * pitch + x;
    const int xr = (x < w - 1) ? p + 1 : p, xl = (x > 0) ? p - 1 : p;
    const int yd = (y < h - 1) ? p + pitch : p, yu = (y > 0) ? p - pitch : p;
    const float *J = occ_plane(a, OC_IM1, b);
    occ_plane(a, OC_JX, b)[p] = (float)(0.5 * (J[xr] - J[xl]));
    occ_plane(a, OC_JY, b)[p] = (float)(0.5 * (J[yd] - J[yu]));
    const float *I0 = a.I0 + (size_t)b * a.g.plane;
    const float i0x = (float)(0.5 * (I0[xr] - I0[xl])), i0y = (float)(0.5 * (I0[yd] - I0[yu]));
    const float gr = sqrtf(i0x * i0x + i0y * i0y);
    const float gamma = 0.05f;
    occ_plane(a, OC_G, b)[p] = 1 / (1 + gamma * gr);
    occ_plane(a, OC_UBA1, b)[p] = -occ_plane(a, OC_U1, b)[p];
    occ_plane(a, OC_UBA2, b)[p] = -occ_plane(a, OC_U2, b)[p];
#pragma unroll
    for (int k = 0; k < 4; k++) occ_plane(a, OC_XI0 + k, b)[p] = 0.f;
}

// per warp: 6 bicubic warps with border_out = false (:616-623) + constants (:628-648)
__global__ void __launch_bounds__(256) occ_warp_kernel(OccArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    const int w = a.g.w, h = a.g.h, pitch = a.g.pitch;
    if (x >= w || y >= h) return;
    const int p = y * pitch + x;
    const size_t off = (size_t)b * a.g.plane;
    const float u1 = occ_plane(a, OC_U1, b)[p], u2 = occ_plane(a, OC_U2, b)[p];
    const float i0 = a.I0[off + p];
    {
        const Taps t = make_taps(x + u1, y + u2, w, h);
        const float iw = bicubic_sample(a.I1 + off, t, pitch);
        const float ix = bicubic_sample(a.I1x + off, t, pitch);
        const float iy = bicubic_sample(a.I1y + off, t, pitch);
        occ_plane(a, OC_I1WX, b)[p] = ix;
        occ_plane(a, OC_I1WY, b)[p] = iy;
        occ_plane(a, OC_RHO_C1, b)[p] = iw - ix * u1 - iy * u2 - i0;
    }
    {
        const float ub1 = occ_plane(a, OC_UBA1, b)[p], ub2 = occ_plane(a, OC_UBA2, b)[p];
        const Taps t = make_taps(x + ub1, y + ub2, w, h);
        const float jw = bicubic_sample(occ_plane(a, OC_IM1, b), t, pitch);
        const float jx = bicubic_sample(occ_plane(a, OC_JX, b), t, pitch);
        const float jy = bicubic_sample(occ_plane(a, OC_JY, b), t, pitch);
        occ_plane(a, OC_JWX, b)[p] = jx;
        occ_plane(a, OC_JWY, b)[p] = jy;
        occ_plane(a, OC_RHO_C_1, b)[p] = jw - jx * u1 - jy * u2 - i0;
    }
}

// v-update with the chi switch (:657-713) + k = theta*beta*grad(chi) (:716, used at :353-354 and :736-737)
__global__ void __launch_bounds__(256) occ_v_kernel(OccArgs a, int it) {
    const int b = blockIdx.z;
    if (!occ_active(a, b, it)) return;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int w = a.g.w, h = a.g.h, pitch = a.g.pitch;
    if (x >= w || y >= h) return;
    const int p = y * pitch + x;
    const float u1 = occ_plane(a, OC_U1, b)[p], u2 = occ_plane(a, OC_U2, b)[p];
    const float *chi = occ_plane(a, OC_CHI0, b);
    const float c = chi[p];
    const float I1wx = occ_plane(a, OC_I1WX, b)[p], I1wy = occ_plane(a, OC_I1WY, b)[p];
    const float Jwx = occ_plane(a, OC_JWX, b)[p], Jwy = occ_plane(a, OC_JWY, b)[p];
    const float rho_1 = occ_plane(a, OC_RHO_C1, b)[p] + I1wx * u1 + I1wy * u2;
    const float rho__1 = occ_plane(a, OC_RHO_C_1, b)[p] + Jwx * u1 + Jwy * u2;
    const float alpha = a.alpha, theta = a.theta, l_t = a.l_t;
    int eps;
    float alpha_i, mu, Lambda, grad, Iwx, Iwy, rho;
    if (c == 0) {
        eps = 1;
        alpha_i = 1;
        mu = l_t;
        Lambda = rho_1;
        grad = I1wx * I1wx + I1wy * I1wy;
        Iwx = I1wx;
        Iwy = I1wy;
        rho = rho_1;
    } else {
        eps = -1;
        alpha_i = 1 / (1 + alpha * theta);
        mu = l_t / (1 + alpha * theta);
        Lambda = rho__1 + alpha * theta / (1 + alpha * theta) * (u1 * Jwx + u2 * Jwy);
        grad = Jwx * Jwx + Jwy * Jwy;
        Iwx = Jwx;
        Iwy = Jwy;
        rho = rho__1;
    }
    float v1, v2;
    if (Lambda > mu * grad) {
        v1 = alpha_i * u1 - mu * eps * Iwx;
        v2 = alpha_i * u2 - mu * eps * Iwy;
    } else if (Lambda < -mu * grad) {
        v1 = alpha_i * u1 + mu * eps * Iwx;
        v2 = alpha_i * u2 + mu * eps * Iwy;
    } else if (grad_is_zero(grad)) {
        v1 = u1;
        v2 = u2;
    } else {
        v1 = u1 - eps * rho * Iwx / grad;
        v2 = u2 - eps * rho * Iwy / grad;
    }
    occ_plane(a, OC_V1, b)[p] = v1;
    occ_plane(a, OC_V2, b)[p] = v2;
    const float chix = (x < w - 1) ? chi[p + 1] - c : 0.f;
    const float chiy = (y < h - 1) ? chi[p + pitch] - c : 0.f;
    occ_plane(a, OC_K1, b)[p] = theta * a.beta * chix;
    occ_plane(a, OC_K2, b)[p] = theta * a.beta * chiy;
}

// div(g*xi_a, g*xi_b) at pixel (x,y), recomputing the products (:343-351)
__device__ __forceinline__ float occ_div_gxi(const float *__restrict__ g, const float *__restrict__ xa,
                                             const float *__restrict__ xb, int x, int y, int w, int h, int pitch) {
    const int p = y * pitch + x;
    const float a_c = g[p] * xa[p];
    const float b_c = g[p] * xb[p];
    const float a_l = (x > 0) ? g[p - 1] * xa[p - 1] : 0.f;
    const float b_u = (y > 0) ? g[p - pitch] * xb[p - pitch] : 0.f;
    return div_bc(a_c, a_l, b_c, b_u, x, y, w, h);
}

// final divergence, u-update, |du|^2, F and G (:394-406, :726-751)
__global__ void __launch_bounds__(256) occ_u_kernel(OccArgs a, int it) {
    const int b = blockIdx.z;
    if (!occ_active(a, b, it)) return;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int w = a.g.w, h = a.g.h, pitch = a.g.pitch;
    float e = 0.f;
    if (x < w && y < h) {
        const int p = y * pitch + x;
        const float *g = occ_plane(a, OC_G, b);
        const float *xi = occ_plane(a, OC_XI0, b);  // 24 sweeps: back in set 0
        const size_t ks = (size_t)a.g.B * a.g.plane;
        const float theta = a.theta;
        const float v1 = occ_plane(a, OC_V1, b)[p], v2 = occ_plane(a, OC_V2, b)[p];
        const float dv1 = occ_div_gxi(g, xi, xi + ks, x, y, w, h, pitch);
        const float dv2 = occ_div_gxi(g, xi + 2 * ks, xi + 3 * ks, x, y, w, h, pitch);
        float *U1 = occ_plane(a, OC_U1, b), *U2 = occ_plane(a, OC_U2, b);
        const float u1k = U1[p], u2k = U2[p];
        const float u1 = v1 + theta * dv1 + occ_plane(a, OC_K1, b)[p];
        const float u2 = v2 + theta * dv2 + occ_plane(a, OC_K2, b)[p];
        U1[p] = u1;
        U2[p] = u2;
        e = (u1 - u1k) * (u1 - u1k) + (u2 - u2k) * (u2 - u2k);
        const float rho__1 = occ_plane(a, OC_RHO_C_1, b)[p] + occ_plane(a, OC_JWX, b)[p] * v1 + occ_plane(a, OC_JWY, b)[p] * v2;
        const float rho_1 = occ_plane(a, OC_RHO_C1, b)[p] + occ_plane(a, OC_I1WX, b)[p] * v1 + occ_plane(a, OC_I1WY, b)[p] * v2;
        occ_plane(a, OC_F, b)[p] = a.lambda * (fabsf(rho__1) - fabsf(rho_1));
        occ_plane(a, OC_GG, b)[p] = a.alpha / 2 * (v1 * v1 + v2 * v2);
    }
    __shared__ float red[8];
    e = warp_max(e);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if ((tid & 31) == 0) red[tid >> 5] = e;
    __syncthreads();
    if (tid == 0) {
        float m = red[0];
        for (int i = 1; i < (int)(blockDim.x * blockDim.y / 32); i++) m = fmaxf(m, red[i]);
        atomicMax(a.err_max + (size_t)b * a.max_iters + it, __float_as_uint(m));
    }
}

// ---------------------------------------------------------------------------
// Register-resident variant of the temporally blocked sweeps (the default).
// One WARP owns one staged row, one LANE one float4 quad of it, for all NS sweeps: the
// lane's own xi / eta / chi and the constants g, v, k (F, G) live in registers; only what a
// neighbour needs crosses lanes -- left / right neighbours by warp shuffle, the rows above /
// below through two shared-memory planes.  Staged region: 128 columns (x0-4 .. x0+123) by
// OR_TH + 2 NS rows; written back: the inner 120 x OR_TH pixels.  NS <= 4 (the column apron).
// ---------------------------------------------------------------------------
#ifndef FALDOI_OCC_TH
#define FALDOI_OCC_TH 16
#endif
enum { OR_TH = FALDOI_OCC_TH, OR_W = 120, OR_PW = 128 };

__device__ __forceinline__ void q2a(const float4 &q, float (&x)[4]) { x[0] = q.x, x[1] = q.y, x[2] = q.z, x[3] = q.w; }
__device__ __forceinline__ float4 a2q(const float (&x)[4]) { return make_float4(x[0], x[1], x[2], x[3]); }

template <int NS>
__global__ void __launch_bounds__(32 * (OR_TH + 2 * NS)) occ_xi_rows_kernel(OccArgs a, int it, int src) {
    constexpr int ROWS = OR_TH + 2 * NS;
    __shared__ __align__(16) float sB[2][ROWS][OR_PW];  // g*xi12, g*xi22 of every staged row (for the row below)
    __shared__ __align__(16) float sV[2][ROWS][OR_PW];  // vi1, vi2 (for the row above)
    const int b = blockIdx.z;
    if (!occ_active(a, b, it)) return;
    const int w = a.g.w, h = a.g.h, pitch = a.g.pitch;
    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int x0 = blockIdx.x * OR_W, y0 = blockIdx.y * OR_TH;
    const int gx0 = x0 - 4 + 4 * lane, gy = y0 - NS + r;
    const int cx = 4 * lane;
    const size_t ks = (size_t)a.g.B * a.g.plane;
    const float theta = a.theta, tt = a.tau_theta;
    const bool rowin = gy >= 0 && gy < h;
    const bool in = rowin && gx0 >= 0 && gx0 < pitch;
    const size_t p = (size_t)gy * pitch + gx0;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *xin = occ_plane(a, src ? OC_XI1 : OC_XI0, b);
    float x11[4], x12[4], x21[4], x22[4], g[4], v1[4], v2[4], k1[4], k2[4];
    q2a(in ? ld4(xin + p) : z4, x11);
    q2a(in ? ld4(xin + ks + p) : z4, x12);
    q2a(in ? ld4(xin + 2 * ks + p) : z4, x21);
    q2a(in ? ld4(xin + 3 * ks + p) : z4, x22);
    q2a(in ? ld4(occ_plane(a, OC_G, b) + p) : z4, g);
    q2a(in ? ld4(occ_plane(a, OC_V1, b) + p) : z4, v1);
    q2a(in ? ld4(occ_plane(a, OC_V2, b) + p) : z4, v2);
    q2a(in ? ld4(occ_plane(a, OC_K1, b) + p) : z4, k1);
    q2a(in ? ld4(occ_plane(a, OC_K2, b) + p) : z4, k2);

#pragma unroll 1
    for (int s = 0; s < NS; s++) {
        float a1[4], b1[4], a2[4], b2[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            a1[k] = g[k] * x11[k];
            b1[k] = g[k] * x12[k];
            a2[k] = g[k] * x21[k];
            b2[k] = g[k] * x22[k];
        }
        *reinterpret_cast<float4 *>(&sB[0][r][cx]) = a2q(b1);
        *reinterpret_cast<float4 *>(&sB[1][r][cx]) = a2q(b2);
        __syncthreads();
        float ub1[4], ub2[4];
        q2a(r > 0 ? *reinterpret_cast<const float4 *>(&sB[0][r - 1][cx]) : z4, ub1);
        q2a(r > 0 ? *reinterpret_cast<const float4 *>(&sB[1][r - 1][cx]) : z4, ub2);
        const float la1 = __shfl_up_sync(0xffffffffu, a1[3], 1), la2 = __shfl_up_sync(0xffffffffu, a2[3], 1);
        // Temporal blocking shrinks the region that still matters by one row per sweep (a trapezoid):
        // after sweep s only rows s+1 .. ROWS-2-s hold values the tile's result depends on, so whole
        // warps (warp = row) skip the arithmetic of the other rows and of rows outside the frame.
        const bool row_vi = rowin && r > s && r < ROWS - s;
        const bool row_x = rowin && r > s && r < ROWS - 1 - s;
        float vi1[4] = {0.f, 0.f, 0.f, 0.f}, vi2[4] = {0.f, 0.f, 0.f, 0.f};
        if (row_vi) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int gx = gx0 + k;
                vi1[k] = v1[k] + theta * div_bc(a1[k], k ? a1[k - 1] : la1, b1[k], ub1[k], gx, gy, w, h) + k1[k];
                vi2[k] = v2[k] + theta * div_bc(a2[k], k ? a2[k - 1] : la2, b2[k], ub2[k], gx, gy, w, h) + k2[k];
            }
            *reinterpret_cast<float4 *>(&sV[0][r][cx]) = a2q(vi1);
            *reinterpret_cast<float4 *>(&sV[1][r][cx]) = a2q(vi2);
        }
        __syncthreads();
        if (!row_x) continue;  // (warp-uniform; the barriers of the next sweep are still reached)
        float dn1[4], dn2[4];
        q2a(r < ROWS - 1 ? *reinterpret_cast<const float4 *>(&sV[0][r + 1][cx]) : z4, dn1);
        q2a(r < ROWS - 1 ? *reinterpret_cast<const float4 *>(&sV[1][r + 1][cx]) : z4, dn2);
        const float rv1 = __shfl_down_sync(0xffffffffu, vi1[0], 1), rv2 = __shfl_down_sync(0xffffffffu, vi2[0], 1);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int gx = gx0 + k;
            {
                const float gxv = (gx < w - 1) ? (k < 3 ? vi1[k + 1] : rv1) - vi1[k] : 0.f;
                const float gyv = (gy < h - 1) ? dn1[k] - vi1[k] : 0.f;
                const float e1 = g[k] * gxv, e2 = g[k] * gyv;
                const float nrm = sqrt_or_zero(e1 * e1 + e2 * e2);
                float q1 = x11[k] + tt * e1, q2 = x12[k] + tt * e2;
                div2_shared(q1, q2, 1 + tt * nrm);
                x11[k] = q1;
                x12[k] = q2;
            }
            {
                const float gxv = (gx < w - 1) ? (k < 3 ? vi2[k + 1] : rv2) - vi2[k] : 0.f;
                const float gyv = (gy < h - 1) ? dn2[k] - vi2[k] : 0.f;
                const float e1 = g[k] * gxv, e2 = g[k] * gyv;
                const float nrm = sqrt_or_zero(e1 * e1 + e2 * e2);
                float q1 = x21[k] + tt * e1, q2 = x22[k] + tt * e2;
                div2_shared(q1, q2, 1 + tt * nrm);
                x21[k] = q1;
                x22[k] = q2;
            }
        }
    }
    if (in && r >= NS && r < NS + OR_TH && lane >= 1 && lane <= OR_W / 4) {
        float *xout = occ_plane(a, src ? OC_XI0 : OC_XI1, b);
        st4(xout + p, a2q(x11));
        st4(xout + ks + p, a2q(x12));
        st4(xout + 2 * ks + p, a2q(x21));
        st4(xout + 3 * ks + p, a2q(x22));
    }
}

template <int NS>
__global__ void __launch_bounds__(32 * (OR_TH + 2 * NS)) occ_chi_rows_kernel(OccArgs a, int it, int src, int last) {
    constexpr int ROWS = OR_TH + 2 * NS;
    __shared__ __align__(16) float sC[ROWS][OR_PW];  // chi of every staged row (for the row above)
    __shared__ __align__(16) float sE[ROWS][OR_PW];  // g*eta2 (for the row below)
    const int b = blockIdx.z;
    if (!occ_active(a, b, it)) return;
    const int w = a.g.w, h = a.g.h, pitch = a.g.pitch;
    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int x0 = blockIdx.x * OR_W, y0 = blockIdx.y * OR_TH;
    const int gx0 = x0 - 4 + 4 * lane, gy = y0 - NS + r;
    const int cx = 4 * lane;
    const size_t ks = (size_t)a.g.B * a.g.plane;
    const float mte = a.mu * a.tau_eta;
    const bool rowin = gy >= 0 && gy < h;
    const bool in = rowin && gx0 >= 0 && gx0 < pitch;
    const size_t p = (size_t)gy * pitch + gx0;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *e_in = occ_plane(a, src ? OC_ETA1 : OC_ETA0, b);
    float e1[4], e2[4], c[4], g[4], F[4], G[4];
    q2a(in ? ld4(e_in + p) : z4, e1);
    q2a(in ? ld4(e_in + ks + p) : z4, e2);
    q2a(in ? ld4(occ_plane(a, src ? OC_CHI1 : OC_CHI0, b) + p) : z4, c);
    q2a(in ? ld4(occ_plane(a, OC_G, b) + p) : z4, g);
    q2a(in ? ld4(occ_plane(a, OC_F, b) + p) : z4, F);
    q2a(in ? ld4(occ_plane(a, OC_GG, b) + p) : z4, G);

#pragma unroll 1
    for (int s = 0; s < NS; s++) {
        *reinterpret_cast<float4 *>(&sC[r][cx]) = a2q(c);
        __syncthreads();
        float dn[4];
        q2a(r < ROWS - 1 ? *reinterpret_cast<const float4 *>(&sC[r + 1][cx]) : z4, dn);
        const float rc = __shfl_down_sync(0xffffffffu, c[0], 1);
        // trapezoid (see occ_xi_rows_kernel): eta matters on rows s .. ROWS-2-s, chi on s+1 .. ROWS-2-s
        const bool row_e = rowin && r >= s && r < ROWS - 1 - s;
        const bool row_c = rowin && r > s && r < ROWS - 1 - s;
        float ge1[4] = {0.f, 0.f, 0.f, 0.f}, ge2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (!row_e) break;
            const int gx = gx0 + k;
            const float chix = (gx < w - 1) ? (k < 3 ? c[k + 1] : rc) - c[k] : 0.f;
            const float chiy = (gy < h - 1) ? dn[k] - c[k] : 0.f;
            const float n1 = e1[k] + mte * g[k] * chix;
            const float n2 = e2[k] + mte * g[k] * chiy;
            const float ne = sqrt_or_zero(n1 * n1 + n2 * n2);
            if (ne <= 1) {
                e1[k] = n1;
                e2[k] = n2;
            } else {
                e1[k] = n1 / ne;
                e2[k] = n2 / ne;
            }
            ge1[k] = g[k] * e1[k];
            ge2[k] = g[k] * e2[k];
        }
        *reinterpret_cast<float4 *>(&sE[r][cx]) = a2q(ge2);
        __syncthreads();
        if (!row_c) continue;  // (warp-uniform)
        float up[4];
        q2a(r > 0 ? *reinterpret_cast<const float4 *>(&sE[r - 1][cx]) : z4, up);
        const float le = __shfl_up_sync(0xffffffffu, ge1[3], 1);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int gx = gx0 + k;
            const float dge = div_bc(ge1[k], k ? ge1[k - 1] : le, ge2[k], up[k], gx, gy, w, h);
            const float div_u = 0.f;  // target definition
            const float cn = c[k] + a.tau_chi * (a.mu * dge - a.beta * div_u - F[k] - G[k]);
            const float lo = (cn < 1) ? cn : 1;
            c[k] = (lo > 0) ? lo : 0;
        }
    }
    if (in && r >= NS && r < NS + OR_TH && lane >= 1 && lane <= OR_W / 4) {
        float *e_out = occ_plane(a, src ? OC_ETA0 : OC_ETA1, b);
        st4(e_out + p, a2q(e1));
        st4(e_out + ks + p, a2q(e2));
        if (last) {
#pragma unroll
            for (int k = 0; k < 4; k++) c[k] = ((double)c[k] > 0.6) ? 1.f : 0.f;
        }
        st4(occ_plane(a, src ? OC_CHI0 : OC_CHI1, b) + p, a2q(c));
    }
}

__global__ void occ_export_kernel(OccArgs a, float *packed) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= a.g.w || y >= a.g.h) return;
    const size_t n = (size_t)a.g.w * a.g.h;
    const int p = y * a.g.pitch + x;
    float *o = packed + (size_t)b * 3 * n + (size_t)y * a.g.w + x;
    o[0] = occ_plane(a, OC_U1, b)[p];
    o[n] = occ_plane(a, OC_U2, b)[p];
    o[2 * n] = occ_plane(a, OC_CHI0, b)[p];
}

}  // namespace faldoi
