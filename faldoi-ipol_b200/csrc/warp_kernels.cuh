// Per-run and per-warp kernels: centred gradient, bicubic warp of I1 and its
// gradients fused with the per-warp constants of each energy model.
#pragma once
#include "common.cuh"

namespace faldoi {

// ---------------------------------------------------------------------------
// centered_gradient (src/utils.cpp:367-423): 0.5*(f[+1]-f[-1]), one-sided
// 0.5*(f[1]-f[0]) at the borders.  The float difference is halved in double in
// the reference; halving is exact, so 0.5f* gives the same bits.
// ---------------------------------------------------------------------------
__global__ void centered_gradient_kernel(const float *__restrict__ f, float *__restrict__ dx,
                                         float *__restrict__ dy, Geo g) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.w || y >= g.h) return;
    const size_t off = (size_t)blockIdx.z * g.plane;
    const float *F = f + off;
    const int p = y * g.pitch + x;
    const int xr = (x < g.w - 1) ? p + 1 : p, xl = (x > 0) ? p - 1 : p;
    const int yd = (y < g.h - 1) ? p + g.pitch : p, yu = (y > 0) ? p - g.pitch : p;
    dx[off + p] = (float)(0.5 * (F[xr] - F[xl]));
    dy[off + p] = (float)(0.5 * (F[yd] - F[yu]));
}

// Keys cubic in one dimension, evaluated in double exactly as
// cubic_interpolation_cell (src/bicubic_interpolation.c:103-111).
__device__ __forceinline__ float keys_cell(float v0, float v1, float v2, float v3, float t) {
    return (float)(v1 + 0.5 * t * (v2 - v0 + t * (2.0 * v0 - 5.0 * v1 + 4.0 * v2 - v3 + t * (3.0 * (v1 - v2) + v3 - v0))));
}

// Tap geometry of bicubic_interpolation_at (src/bicubic_interpolation.c:138-237):
// truncation toward zero, taps x-sx, x, x+sx, x+2sx, Neumann clamp, and the
// row above uses sx instead of sy (:159).  `out` = some index was clamped.
struct Taps {
    int cx[4], cy[4];
    float tx, ty;
    bool out;
};

__device__ __forceinline__ int clamp_idx(int v, int n, bool &out) {
    if (v < 0) {
        out = true;
        return 0;
    }
    if (v >= n) {
        out = true;
        return n - 1;
    }
    return v;
}

__device__ __forceinline__ Taps make_taps(float uu, float vv, int w, int h) {
    Taps t;
    const int sx = (uu < 0) ? -1 : 1, sy = (vv < 0) ? -1 : 1;
    const int xi = (int)uu, yi = (int)vv;
    t.out = false;
    t.cx[1] = clamp_idx(xi, w, t.out);
    t.cy[1] = clamp_idx(yi, h, t.out);
    t.cx[0] = clamp_idx(xi - sx, w, t.out);
    t.cy[0] = clamp_idx(yi - sx, h, t.out);
    t.cx[2] = clamp_idx(xi + sx, w, t.out);
    t.cy[2] = clamp_idx(yi + sy, h, t.out);
    t.cx[3] = clamp_idx(xi + 2 * sx, w, t.out);
    t.cy[3] = clamp_idx(yi + 2 * sy, h, t.out);
    t.tx = uu - t.cx[1];
    t.ty = vv - t.cy[1];
    return t;
}

__device__ __forceinline__ float bicubic_sample(const float *__restrict__ img, const Taps &t, int pitch) {
    float c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float *col = img + t.cx[k];
        c[k] = keys_cell(__ldg(col + t.cy[0] * pitch), __ldg(col + t.cy[1] * pitch), __ldg(col + t.cy[2] * pitch),
                         __ldg(col + t.cy[3] * pitch), t.ty);
    }
    return keys_cell(c[0], c[1], c[2], c[3], t.tx);
}

// ---------------------------------------------------------------------------
// Per-warp kernel of the TV-L1 data term (methods 0-3): warp I1, I1x, I1y with
// the current flow (border_out = true) and build the constants of
// src/global_faldoi.cpp:635-660:  Ix, Iy, rho_c = I1w - Ix*u1 - Iy*u2 - I0 and
// ubar = u.  grad = Ix^2+Iy^2 is recomputed by the iteration kernel (same bits
// without FMA) instead of being stored.
// When I1w_out != nullptr (CSAD methods) the warped image is stored instead of
// rho_c (the CSAD constants kernel consumes it).
// ---------------------------------------------------------------------------
struct WarpArgs {
    const float *I0, *I1, *I1x, *I1y;  // static planes [B]; in stripe mode I1* are FULL frames (any row can be sampled)
    size_t i1_plane;                   // plane stride of I1, I1x, I1y
    const float *u1, *u2;              // base of the two flow planes (pair stride = plane)
    float *ub1, *ub2;                  // extrapolated flow, set to u
    float *Ix, *Iy, *rho_c, *I1w;      // outputs [B]
    const int *parity;                 // per pair: which ping-pong set is current
    size_t set_stride;                 // floats between ping-pong set 0 and set 1
    Geo g;
};

__global__ void __launch_bounds__(256) warp_constants_kernel(WarpArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= a.g.w || y >= a.g.h) return;
    const size_t off = (size_t)b * a.g.plane;
    const size_t so = off + (size_t)a.parity[b] * a.set_stride;
    const int p = y * a.g.pitch + x;
    const float u1 = a.u1[so + p], u2 = a.u2[so + p];
    const float uu = x + u1, vv = (y + a.g.y_off) + u2;
    const Taps t = make_taps(uu, vv, a.g.w, a.g.hg);
    float iw = 0.f, ix = 0.f, iy = 0.f;
    if (!t.out) {
        const size_t off1 = (size_t)b * a.i1_plane;
        iw = bicubic_sample(a.I1 + off1, t, a.g.pitch);
        ix = bicubic_sample(a.I1x + off1, t, a.g.pitch);
        iy = bicubic_sample(a.I1y + off1, t, a.g.pitch);
    }
    a.Ix[off + p] = ix;
    a.Iy[off + p] = iy;
    if (a.I1w) a.I1w[off + p] = iw;
    if (a.rho_c) a.rho_c[off + p] = iw - ix * u1 - iy * u2 - a.I0[off + p];
    a.ub1[so + p] = u1;
    a.ub2[so + p] = u2;
}

// Plain bicubic warp of one image (src/bicubic_interpolation.c:245-266), used by
// the occlusion model (border_out = false) and the standalone C-ABI helper.
__global__ void __launch_bounds__(256) bicubic_warp_kernel(const float *__restrict__ img, const float *__restrict__ u,
                                                           const float *__restrict__ v, float *__restrict__ out,
                                                           float flow_sign, int border_out, Geo g) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.w || y >= g.h) return;
    const size_t off = (size_t)blockIdx.z * g.plane;
    const int p = y * g.pitch + x;
    const float uu = x + flow_sign * u[off + p], vv = y + flow_sign * v[off + p];
    const Taps t = make_taps(uu, vv, g.w, g.h);
    out[off + p] = (t.out && border_out) ? 0.f : bicubic_sample(img + off, t, g.pitch);
}

}  // namespace faldoi
