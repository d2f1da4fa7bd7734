#include "preprocess.h"

#include <cmath>
#include <cstring>
#include <stdexcept>

namespace faldoi_host {

void rgb2gray(const float *rgb, int w, int h, float *out) {
    const size_t n = (size_t)w * h;
    const float *r = rgb, *g = rgb + n, *b = rgb + 2 * n;
    for (size_t i = 0; i < n; i++) out[i] = static_cast<float>(.299 * r[i] + .587 * g[i] + .114 * b[i]);
}

namespace {
struct Range {
    float lo, hi;
};
Range range_of(const float *x, int n) {
    Range r{x[0], x[0]};
    for (int i = 1; i < n; i++) {
        if (x[i] < r.lo) r.lo = x[i];
        if (x[i] > r.hi) r.hi = x[i];
    }
    return r;
}
}  // namespace

// The reference is called as image_normalization_3(i0, i1, i_1, ...), whose formals are
// (I1, I2, I0): the joint maximum is the true maximum, but the joint "minimum" is
// max(min(i1), min(min(i_1), min(i0))) (src/utils.cpp:763 picks the larger one).
void normalize3(float *i0, float *i1, float *im1, int n) {
    const Range a = range_of(im1, n), b = range_of(i0, n), c = range_of(i1, n);
    const float hi_ab = a.hi > b.hi ? a.hi : b.hi;
    const float hi = c.hi > hi_ab ? c.hi : hi_ab;
    const float lo_ab = a.lo < b.lo ? a.lo : b.lo;
    const float lo = c.lo > lo_ab ? c.lo : lo_ab;
    const float den = hi - lo;
    if (!(den > 0)) return;  // constant images are copied through
    for (int i = 0; i < n; i++) {
        im1[i] = (im1[i] - lo) / den;
        i0[i] = (i0[i] - lo) / den;
        i1[i] = (i1[i] - lo) / den;
    }
}

void gaussian(float *img, int w, int h, float sigma) {
    const float den = 2 * sigma * sigma;
    const int radius = static_cast<int>(5 * sigma);  // the reference's `size` is radius + 1
    const int taps = radius + 1;
    if (taps > w || taps > h) throw std::runtime_error("gaussian: sigma too large for the image");
    std::vector<float> k(taps);
    for (int i = 0; i < taps; i++)
        k[i] = static_cast<float>(1 / (sigma * std::sqrt(2.0 * 3.1415926)) * std::exp(static_cast<float>(-i * i) / den));
    float norm = 0;
    for (float v : k) norm += v;
    norm *= 2;
    norm -= k[0];
    for (float &v : k) v /= norm;

    // one padded line buffer: [taps pad | samples | taps pad].  The low-side pad mirrors
    // WITHOUT repeating the edge sample (line[i] = s[taps - i]), the high-side pad mirrors
    // WITH it (src/utils.cpp:569-573) -- asymmetric on purpose, it is what the reference does.
    std::vector<float> line(static_cast<size_t>(w > h ? w : h) + 2 * taps);
    auto filter_line = [&](float *first, size_t stride, int len) {
        for (int i = 0; i < len; i++) line[taps + i] = first[i * stride];
        for (int i = 0; i < taps; i++) {
            line[i] = first[(taps - i) * stride];
            line[taps + len + i] = first[(len - i - 1) * stride];
        }
        for (int i = 0; i < len; i++) {
            const float *c = &line[taps + i];
            float acc = k[0] * c[0];
            for (int j = 1; j < taps; j++) acc += k[j] * (c[-j] + c[j]);
            first[i * stride] = acc;
        }
    };
    for (int y = 0; y < h; y++) filter_line(img + static_cast<size_t>(y) * w, 1, w);
    for (int x = 0; x < w; x++) filter_line(img + x, w, h);
}

void image_to_lab(const float *rgb, int n, float *lab) {
    const float T = 0.008856;
    const float attenuation = 1.5f;
    auto sq = [](float f) { return f * f; };
    for (int i = 0; i < n; i++) {
        const float r = rgb[i] / 255.f, g = rgb[i + n] / 255.f, b = rgb[i + 2 * n] / 255.f;
        float X = static_cast<float>(0.412453 * r + 0.357580 * g + 0.180423 * b);
        const float Y = static_cast<float>(0.212671 * r + 0.715160 * g + 0.072169 * b);
        float Z = static_cast<float>(0.019334 * r + 0.119193 * g + 0.950227 * b);
        X = static_cast<float>(X / 0.950456);
        Z = static_cast<float>(Z / 1.088754);
        const float Y3 = static_cast<float>(std::pow(static_cast<double>(Y), 1. / 3));
        auto f_of = [&](float t) {
            return static_cast<float>(t > T ? std::pow(static_cast<double>(t), 1. / 3) : 7.787 * t + 16 / 116.);
        };
        const float fX = f_of(X);
        const float fY = static_cast<float>(Y > T ? static_cast<double>(Y3) : 7.787 * Y + 16 / 116.);
        const float fZ = f_of(Z);
        const float L = static_cast<float>(Y > T ? 116 * Y3 - 16.0 : 903.3 * Y);
        const float A = 500 * (fX - fY);
        const float B = 200 * (fY - fZ);
        // dark / very light areas have less reliable colour: attenuate a,b
        const float corr = std::exp(-attenuation * sq(static_cast<float>(sq(L / 100) - 0.6)));
        lab[i] = L;
        lab[i + n] = A * corr;
        lab[i + 2 * n] = B * corr;
    }
}

void preprocess(const float *i0, const float *i1, const float *im1, int pd, int w, int h, float *i0n, float *i1n,
                float *im1n) {
    const size_t n = static_cast<size_t>(w) * h;
    if (pd != 1) {
        rgb2gray(i0, w, h, i0n);
        rgb2gray(i1, w, h, i1n);
        rgb2gray(im1, w, h, im1n);
    } else {
        std::memcpy(i0n, i0, n * sizeof(float));
        std::memcpy(i1n, i1, n * sizeof(float));
        std::memcpy(im1n, im1, n * sizeof(float));
    }
    normalize3(i0n, i1n, im1n, static_cast<int>(n));
    const float sigma = 0.90f;  // PRESMOOTHING_SIGMA (src/parameters.h:16)
    gaussian(i0n, w, h, sigma);
    gaussian(i1n, w, h, sigma);
    gaussian(im1n, w, h, sigma);
}

}  // namespace faldoi_host
