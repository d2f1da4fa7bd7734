// extern "C" surface of the host-side code (include/faldoi_host.h).
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>

#include "../../include/faldoi_host.h"
#include "image_io.h"
#include "preprocess.h"

static thread_local std::string g_host_err;

template <class F>
static int guarded(F f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_host_err = e.what();
        return 1;
    }
}

extern "C" {

const char *faldoi_host_last_error(void) { return g_host_err.c_str(); }

int faldoi_host_preprocess(const float *i0, const float *i1, const float *im1, int pd, int w, int h, float *i0n,
                           float *i1n, float *im1n) {
    return guarded([&] { faldoi_host::preprocess(i0, i1, im1, pd, w, h, i0n, i1n, im1n); });
}

int faldoi_host_image_to_lab(const float *rgb, int w, int h, float *lab) {
    return guarded([&] { faldoi_host::image_to_lab(rgb, w * h, lab); });
}

int faldoi_host_read_image(const char *path, float **data, int *w, int *h, int *pd) {
    return guarded([&] {
        const faldoi_host::Image im = faldoi_host::read_image_split(path);
        *w = im.w;
        *h = im.h;
        *pd = im.pd;
        *data = static_cast<float *>(std::malloc(im.count() * sizeof(float)));
        std::memcpy(*data, im.px(), im.count() * sizeof(float));
    });
}

void faldoi_host_free(void *p) { std::free(p); }

int faldoi_host_write_flo(const char *path, const float *u1, const float *u2, int w, int h) {
    return guarded([&] { faldoi_host::write_flo(path, u1, u2, w, h); });
}

int faldoi_host_write_png_gray8(const char *path, const int *values, int w, int h) {
    return guarded([&] { faldoi_host::write_png_gray8(path, values, w, h); });
}
}
