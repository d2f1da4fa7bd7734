// main()'s preprocessing on the device (the first "next" row):
// rgb2gray (src/global_faldoi.cpp:1820-1827), image_normalization_3 (src/utils.cpp:743-781),
// gaussian (src/utils.cpp:521-630) and image_to_lab (src/global_faldoi.cpp:906-932).
// Gray / normalise / smooth are bit-identical to the host code (same expression types and
// summation order, no FMA); the Lab conversion uses the device's double pow where the reference calls
// glibc's (both round to the same float except when the double results straddle a float rounding
// boundary, ~2^-29 per call) and glibc's expf algorithm for the attenuation factor: on the full-size
// Sintel frame all 1.3 M Lab values equal the host's (tools/lab_check.py).
#pragma once
#include "common.cuh"

namespace faldoi {

// order-preserving float <-> unsigned key, so atomicMin/Max work for any sign
__device__ __forceinline__ unsigned f2key(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// raw planar frame (pd planes of w*h, dense) -> gray plane (pitched); also folds the plane's
// min / max into mm[0], mm[1] (keys).  One launch per frame.
__global__ void __launch_bounds__(256) gray_minmax_kernel(const float *__restrict__ raw, int pd, float *__restrict__ gray,
                                                           unsigned *__restrict__ mm, Geo g) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    unsigned kmin = 0xffffffffu, kmax = 0u;
    if (x < g.w && y < g.h) {
        const size_t n = (size_t)g.w * g.h, i = (size_t)y * g.w + x;
        float v;
        if (pd != 1)
            v = (float)(.299 * raw[i] + .587 * raw[n + i] + .114 * raw[2 * n + i]);
        else
            v = raw[i];
        gray[(size_t)y * g.pitch + x] = v;
        kmin = kmax = f2key(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0) {
        atomicMin(mm, kmin);
        atomicMax(mm + 1, kmax);
    }
}

// joint normalisation with the reference's argument order (I1,I2,I0) := (i0,i1,i_1):
// max = max of the three maxima, "min" = the LARGER of min(i1) and min(min(i_1), min(i0)).
// mm = {min,max} keys of i0, i1, i_1 in that order.
__global__ void __launch_bounds__(256) normalize3_kernel(float *__restrict__ i0, float *__restrict__ i1, float *__restrict__ im1,
                                                          const unsigned *__restrict__ mm, Geo g) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.w || y >= g.h) return;
    const float mn0 = key2f(mm[0]), mx0 = key2f(mm[1]), mn1 = key2f(mm[2]), mx1 = key2f(mm[3]);
    const float mnm = key2f(mm[4]), mxm = key2f(mm[5]);
    const float max01 = (mxm > mx0) ? mxm : mx0;
    const float mx = (mx1 > max01) ? mx1 : max01;
    const float min01 = (mnm < mn0) ? mnm : mn0;
    const float mn = (mn1 > min01) ? mn1 : min01;
    const float den = mx - mn;
    if (!(den > 0)) return;
    const size_t p = (size_t)y * g.pitch + x;
    im1[p] = (im1[p] - mn) / den;
    i0[p] = (i0[p] - mn) / den;
    i1[p] = (i1[p] - mn) / den;
}

struct GaussTaps {
    float k[8];
    int taps;  // radius + 1
};

// one separable pass of the reference Gaussian; `vertical` selects the axis.  Low-side
// reflection skips the edge sample, high-side reflection repeats it (src/utils.cpp:569-573).
__global__ void __launch_bounds__(256) gaussian_pass_kernel(const float *__restrict__ in, float *__restrict__ out, GaussTaps t,
                                                             int vertical, Geo g) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.w || y >= g.h) return;
    const int len = vertical ? g.h : g.w, i = vertical ? y : x;
    const size_t base = vertical ? (size_t)x : (size_t)y * g.pitch;
    const size_t stride = vertical ? (size_t)g.pitch : 1;
    auto at = [&](int idx) {
        if (idx < 0) idx = -idx;
        if (idx >= len) idx = 2 * len - idx - 1;
        return in[base + (size_t)idx * stride];
    };
    float sum = t.k[0] * at(i);
    for (int j = 1; j < t.taps; j++) sum += t.k[j] * (at(i - j) + at(i + j));
    out[(size_t)y * g.pitch + x] = sum;
}

// raw planar rgb (dense, 0..255) -> Lab planes (pitched)
__global__ void __launch_bounds__(256) image_to_lab_kernel(const float *__restrict__ rgb, float *__restrict__ L, float *__restrict__ A,
                                                            float *__restrict__ Bp, Geo g) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.w || y >= g.h) return;
    const size_t n = (size_t)g.w * g.h, i = (size_t)y * g.w + x, p = (size_t)y * g.pitch + x;
    const float T = 0.008856;
    const float r = rgb[i] / 255.f, gg = rgb[n + i] / 255.f, b = rgb[2 * n + i] / 255.f;
    float X = (float)(0.412453 * r + 0.357580 * gg + 0.180423 * b);
    const float Y = (float)(0.212671 * r + 0.715160 * gg + 0.072169 * b);
    float Z = (float)(0.019334 * r + 0.119193 * gg + 0.950227 * b);
    X = (float)(X / 0.950456);
    Z = (float)(Z / 1.088754);
    const float Y3 = (float)pow((double)Y, 1. / 3);
    const float fX = (float)(X > T ? pow((double)X, 1. / 3) : 7.787 * X + 16 / 116.);
    const float fY = (float)(Y > T ? (double)Y3 : 7.787 * Y + 16 / 116.);
    const float fZ = (float)(Z > T ? pow((double)Z, 1. / 3) : 7.787 * Z + 16 / 116.);
    const float Lv = (float)(Y > T ? 116 * Y3 - 16.0 : 903.3 * Y);
    const float Av = 500 * (fX - fY);
    const float Bv = 200 * (fY - fZ);
    const float t0 = (Lv / 100) * (Lv / 100);
    const float t1 = (float)(t0 - 0.6);
    const float corr = expf_glibc_nonpositive(-1.5f * (t1 * t1));  // the reference's exp(float) is expf
    L[p] = Lv;
    A[p] = Av * corr;
    Bp[p] = Bv * corr;
}

}  // namespace faldoi
