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

}  // namespace faldoi
