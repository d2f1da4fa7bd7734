// Host-side preprocessing of the reference's main() (src/global_faldoi.cpp:2042-2068):
// gray conversion, joint normalisation, Gaussian pre-smoothing, Lab conversion.
// These run once per frame on the host in the reference and stay on the host here
// (SURVEY.md 8f lists their GPU version as the next row).
#pragma once
#include <vector>

namespace faldoi_host {

// rgb2gray (src/global_faldoi.cpp:1820-1827): .299 R + .587 G + .114 B, evaluated in double
void rgb2gray(const float *rgb, int w, int h, float *out);
// image_normalization_3 (src/utils.cpp:743-781) with main()'s argument order (:2065), in place
void normalize3(float *i0, float *i1, float *im1, int n);
// gaussian (src/utils.cpp:521-630): separable, in place, sigma -> radius (int)(5 sigma)
void gaussian(float *img, int w, int h, float sigma);
// image_to_lab (src/global_faldoi.cpp:906-932); rgb planar 0..255 -> Lab planar
void image_to_lab(const float *rgb, int n, float *lab);
// the whole sequence; pd = channels of the inputs (1 = already gray)
void preprocess(const float *i0, const float *i1, const float *im1, int pd, int w, int h, float *i0n, float *i1n,
                float *im1n);

}  // namespace faldoi_host
