// NLTV regulariser kernels (methods 2,3: nltvl1_PD src/global_faldoi.cpp:1177-1328;
// methods 6,7: nltvcsad_PD :1642-1808).
//
// Layout: the reference's 388-byte AoS DualVariables_global (sc/wp/ap/rp per
// pixel, :891-897) becomes 24-plane SoA.  Slot s <-> offset (k,l), rows
// k=-2..2 outer, columns l=-2..2 inner, centre skipped, slot counter advancing
// for out-of-image offsets too (:1008-1036); the neighbour index `ap` and the
// reciprocal slot `rp = 23 - s` are implicit.
#pragma once
#include "common.cuh"
#include "tv_kernels.cuh"

namespace faldoi {

enum { NL_SLOTS = 24 };

// Two arithmetic modes (chosen at run time, FALDOI_NLTV_FAST=1 selects the second):
//   exact  every operation of ofnltv_getD / non_local_divergence / ofnltv_getP as the reference writes
//          it: (w*(u_p-u_q))/wt, (sc + tau*g)/(1 + tau*|g|) and div/wt are IEEE divisions, the 24 terms
//          of the divergence are summed in ascending slot order.  With the weights below the NLTV
//          flows are bit-identical to the reference's.
//   fast   a * rcp.approx(b) (MUFU.RCP + FMUL, 2 ulp) for the 192 divisions per pixel and iteration,
//          1/wt formed once per pixel, slots visited as (s, 23-s) pairs.  Inside the north star's mean
//          tolerance; the CSAD variants are chaotic enough (a rank flip moves a pixel by 1e-2 px) that
//          a few pixels of a full-size frame can exceed its max tolerance, which is why it is opt-in.
__device__ __forceinline__ float nl_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#define NL_DIV(a, b) ((a) * nl_rcp(b))

struct NlOffsets {
    float ws[NL_SLOTS];  // exp(-hypot(l,k)/2) per slot, evaluated on the host with libm (get_wspatial_2 :943-952)
};

__host__ __device__ __forceinline__ void nl_slot_offset(int s, int &k, int &l) {
    const int t = s + (s >= 12);  // position in the 5x5 window, centre (12) skipped
    k = t / 5 - 2;
    l = t % 5 - 2;
}

// ---------------------------------------------------------------------------
// initialize_dual_variables (:996-1054): wp = sqrt(w_colour * w_spatial),
// w_colour = exp(-|Lab(p)-Lab(q)|/5) (get_wcolor_2 :954-978), wt = sum of the
// in-image wp in slot order.  expf is glibc's algorithm (expf_glibc_nonpositive above).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nltv_init_kernel(const float *__restrict__ lab, float *__restrict__ wgt,
                                                        float *__restrict__ wt, float *__restrict__ rwt, NlOffsets offs, Geo g) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= g.w || y >= g.h) return;
    const size_t ks = (size_t)g.B * g.plane, off = (size_t)b * g.plane;
    const int p = y * g.pitch + x;
    const float L = lab[off + p], A = lab[ks + off + p], Bc = lab[2 * ks + off + p];
    float ne = 0.f;
#pragma unroll
    for (int s = 0; s < NL_SLOTS; s++) {
        int k, l;
        nl_slot_offset(s, k, l);
        const int r = y + k, c = x + l;
        float wv = -2.0f;
        if (c >= 0 && c < g.w && r >= 0 && r < g.h) {
            const int q = r * g.pitch + c;
            float d = 0.f, aux;
            aux = L - lab[off + q];
            d += aux * aux;
            aux = A - lab[ks + off + q];
            d += aux * aux;
            aux = Bc - lab[2 * ks + off + q];
            d += aux * aux;
            d = sqrtf(d);
            const float wc = expf_glibc_nonpositive(-d / 5.f);
            wv = sqrtf(wc * offs.ws[s]);
            ne += wv;
        }
        wgt[(size_t)s * ks + off + p] = wv;
    }
    wt[off + p] = ne;
    rwt[off + p] = NL_DIV(1.f, ne);  // what the iteration multiplies by (nltv_tile_kernel reads this plane)
}

struct NlArgs {
    float *state;       // [2][ST_COUNT][B]: u1,u2,ub1,ub2 used
    size_t set_stride;
    float *dual;        // [2][2*NL_SLOTS][B]: P slots then Q slots
    size_t dual_set_stride;
    const float *wgt, *wt;   // [24][B], [B]   (+ rwt = 1/wt, read through a tensor map by nltv_tile_kernel)
    const float *Ix, *Iy, *rho_c, *scale, *blk, *sep;  // blk/sep: CSAD table, see csad_select
    double *err_sum;    // [B][max_iters]
    Geo g;
    int max_iters;
    float tau, theta, l_t;
    DivConst dth;  // division by theta
};

// One fused NLTV iteration (:1249-1301 / :1729-1777): data-term v, dual update
// (ofnltv_getD :1127-1174), non-local divergence (:1056-1079) and primal step
// (ofnltv_getP :1090-1120, note +div) in one pass.  The neighbour's reciprocal
// dual P_new[23-s](q) that the divergence needs is recomputed from
// P_old[23-s](q) instead of being re-read after a grid-wide barrier; its
// weight wgt[23-s](q) equals wgt[s](p) bit for bit (symmetric formula).
// Launch `it` reads set (it&1) and writes set (it&1)^1: NLTV always runs all
// max_iters iterations (:1249), so the parity is the same for every pair.
template <int DATA>
__global__ void __launch_bounds__(256, 4) nltv_iter_kernel(NlArgs a, int it, int base_parity) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    const int w = a.g.w, h = a.g.h, pitch = a.g.pitch;
    const bool inimg = (x < w && y < h);
    double esum = 0.0;
    if (inimg) {
        const int par = (base_parity + it) & 1;
        const size_t plane = a.g.plane, ks = (size_t)a.g.B * plane, off = (size_t)b * plane;
        const float *sin = a.state + (size_t)par * a.set_stride + off;
        float *sout = a.state + (size_t)(par ^ 1) * a.set_stride + off;
        const float *din = a.dual + (size_t)par * a.dual_set_stride + off;
        float *dout = a.dual + (size_t)(par ^ 1) * a.dual_set_stride + off;
        const int p = y * pitch + x;
        const float tau = a.tau, l_t = a.l_t;
        const float u1 = sin[ST_U1 * ks + p], u2 = sin[ST_U2 * ks + p];
        const float c1 = sin[ST_UB1 * ks + p], c2 = sin[ST_UB2 * ks + p];
        const float ix = a.Ix[off + p], iy = a.Iy[off + p];
        const float wtp = a.wt[off + p];

        // ---- data term ----
        float v1, v2;
        if (DATA == DATA_TVL1) {
            const float grad = ix * ix + iy * iy;
            const float rho = a.rho_c[off + p] + (ix * u1 + iy * u2);
            float e1, e2;
            if (rho < -l_t * grad) {
                e1 = l_t * ix;
                e2 = l_t * iy;
            } else if (rho > l_t * grad) {
                e1 = -l_t * ix;
                e2 = -l_t * iy;
            } else if (grad_is_zero(grad)) {
                e1 = e2 = 0.f;
            } else {
                const float fi = -rho / grad;
                e1 = fi * ix;
                e2 = fi * iy;
            }
            v1 = u1 + e1;
            v2 = u2 + e2;
        } else {
            v1 = u1;
            v2 = u2;
            const float sc = a.scale[off + p];
            if (sc != 0.f) {  // 0 marks grad <= GRAD_IS_ZERO (:1734)
                const float s = (ix * u1 + iy * u2) / sc;
                const int np = csad_count(x, y, w, h);
                const float med = csad_select(a.blk, a.sep, a.g, b, y, x, np, s, l_t, sc);
                v1 = csad_apply(u1, ix, med, sc);
                v2 = csad_apply(u2, iy, med, sc);
            }
        }

        // ---- dual update + non-local divergence ----
        // Pixels at least 2 away from every frame border (all 24 neighbours exist) take a loop without
        // bounds tests; 1/wt is formed once per pixel (the NLTV models are tolerance-level, see NL_DIV).
        float dP = 0.f, dQ = 0.f;
        const float rwtp = NL_DIV(1.f, wtp);
        const float *ub1p = sin + ST_UB1 * ks, *ub2p = sin + ST_UB2 * ks;
        const float *wtb = a.wt + off;
        auto slot = [&](int s, int q) {
            const float wv = a.wgt[(size_t)s * ks + off + p];
            const float q1 = ub1p[q], q2 = ub2p[q];
            const float rwtq = NL_DIV(1.f, wtb[q]);
            const float t1 = wv * (c1 - q1), t2 = wv * (c2 - q2);
            // own dual, slot s
            const float g1 = t1 * rwtp, g2 = t2 * rwtp;
            const float Pn = NL_DIV(din[(size_t)s * ks + p] + tau * g1, 1 + tau * fabsf(g1));
            const float Qn = NL_DIV(din[(size_t)(NL_SLOTS + s) * ks + p] + tau * g2, 1 + tau * fabsf(g2));
            dout[(size_t)s * ks + p] = Pn;
            dout[(size_t)(NL_SLOTS + s) * ks + p] = Qn;
            // neighbour's reciprocal dual, slot 23-s at q (its difference is the negated one)
            const int rs = NL_SLOTS - 1 - s;
            const float h1 = -t1 * rwtq, h2 = -t2 * rwtq;
            const float Pr = NL_DIV(din[(size_t)rs * ks + q] + tau * h1, 1 + tau * fabsf(h1));
            const float Qr = NL_DIV(din[(size_t)(NL_SLOTS + rs) * ks + q] + tau * h2, 1 + tau * fabsf(h2));
            dP += wv * (Pn - Pr);
            dQ += wv * (Qn - Qr);
        };
        if (x >= 2 && x < w - 2 && y >= 2 && y < h - 2) {
#pragma unroll
            for (int s = 0; s < NL_SLOTS; s++) {
                int k, l;
                nl_slot_offset(s, k, l);
                slot(s, p + k * pitch + l);
            }
        } else {
#pragma unroll
            for (int s = 0; s < NL_SLOTS; s++) {
                int k, l;
                nl_slot_offset(s, k, l);
                const int r = y + k, c = x + l;
                if (c >= 0 && c < w && r >= 0 && r < h) slot(s, r * pitch + c);
            }
        }
        dP *= rwtp;
        dQ *= rwtp;

        // ---- primal step (+div) and extrapolation ----
        const float o1 = u1 - tau * (dP + div_const(u1 - v1, a.dth));
        const float o2 = u2 - tau * (dQ + div_const(u2 - v2, a.dth));
        esum = (double)((o1 - u1) * (o1 - u1) + (o2 - u2) * (o2 - u2));
        sout[ST_U1 * ks + p] = o1;
        sout[ST_U2 * ks + p] = o2;
        sout[ST_UB1 * ks + p] = 2 * o1 - u1;
        sout[ST_UB2 * ks + p] = 2 * o2 - u2;
    }
    // printed error only (the exit test is commented out upstream, :1248)
    __shared__ double red[8];
    esum = warp_sum(esum);
    const int wid = (threadIdx.y * blockDim.x + threadIdx.x) >> 5;
    if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0) red[wid] = esum;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x * blockDim.y / 32); i++) t += red[i];
        atomicAdd(a.err_sum + (size_t)b * a.max_iters + it, t);
    }
}

}  // namespace faldoi
