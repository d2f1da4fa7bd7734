// Image / flow file I/O for the host `global_faldoi` (replaces the reference's use of
// iio: iio_read_image_float_split src/iio.c:2982-2993, .flo read :1807-1823 / write
// :2539-2555, iio_save_image_int :3653-3664).  Own implementation on zlib only:
// PNG (non-interlaced, 1-16 bit, all colour types), binary/ASCII PNM, Middlebury .flo.
#pragma once
#include <string>
#include <vector>

namespace faldoi_host {

// Where the samples of a decoded image go: by default the image's own vector; a caller may hand in storage of its
// own (e.g. a page-locked staging buffer) so that decoding writes straight into it.  Storage that is too small
// for the file is ignored (the image then owns its samples as usual).
struct Sink {
    float *ptr = nullptr;
    size_t cap = 0;  // floats
};

struct Image {
    int w = 0, h = 0, pd = 0;
    std::vector<float> data;  // planar ("split"): data[c*w*h + j*w + i], sample values as stored (0..255 for 8 bit)
    float *ext = nullptr;     // set instead of `data` when the samples were decoded into a caller's Sink
    const float *px() const { return ext ? ext : data.data(); }
    float *px() { return ext ? ext : data.data(); }
    size_t count() const { return (size_t)w * h * pd; }
    bool empty() const { return count() == 0; }
};

// Throws std::runtime_error with a message naming the file on any failure.
Image read_image_split(const std::string &path, Sink sink = Sink());  // dispatch on magic bytes: PNG, PNM, .flo
// Header only: size and channel count as read_image_split would report them.
void probe_image(const std::string &path, int *w, int *h, int *pd);
void write_flo(const std::string &path, const float *u1, const float *u2, int w, int h);
void write_png_gray8(const std::string &path, const int *values, int w, int h);  // values clipped to 0..255
void write_image_float_split(const std::string &path, const float *planes, int w, int h, int pd);  // .flo (pd=2) only

}  // namespace faldoi_host
