/* faldoi_host.h -- C ABI of libfaldoi_host.so: the host-side pieces of the
 * `global_faldoi` executable that surround the GPU solver (file I/O and main()'s
 * preprocessing), exported so other host languages / tests can call them.
 * Pure CPU code; it is preprocessing and I/O, not a solver fallback.
 */
#ifndef FALDOI_HOST_H
#define FALDOI_HOST_H
#ifdef __cplusplus
extern "C" {
#endif

/* main()'s preprocessing (src/global_faldoi.cpp:2049-2068): rgb2gray (pd != 1),
 * image_normalization_3, gaussian(sigma = 0.9).  Inputs planar, pd channels, 0..255. */
int faldoi_host_preprocess(const float *i0, const float *i1, const float *im1, int pd, int w, int h, float *i0n,
                           float *i1n, float *im1n);
/* image_to_lab (src/global_faldoi.cpp:906-932), planar rgb 0..255 -> planar Lab. */
int faldoi_host_image_to_lab(const float *rgb, int w, int h, float *lab);

/* iio_read_image_float_split (src/iio.c:2982-2993) for PNG / PNM / .flo.  On success
 * *data is malloc()ed planar float (free with faldoi_host_free). */
int faldoi_host_read_image(const char *path, float **data, int *w, int *h, int *pd);
void faldoi_host_free(void *p);
/* iio_save_image_float_split(..., 2) for .flo (src/iio.c:2539-2555). */
int faldoi_host_write_flo(const char *path, const float *u1, const float *u2, int w, int h);
/* iio_save_image_int (src/iio.c:3653-3664) as an 8-bit gray PNG. */
int faldoi_host_write_png_gray8(const char *path, const int *values, int w, int h);
const char *faldoi_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
