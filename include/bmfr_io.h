/* bmfr_io.h — dataset ingestion for the driver program: what the reference does with
 * OpenImageIO and an #include of the dataset's camera_matrices.h, rebuilt without either.
 *
 * Host-only C ABI (no CUDA, no torch): libbmfr_io.so, also linked into bmfr_run.  Every function returns
 * 0 on success and a negative BMFR_IO_ERR_* otherwise; bmfr_io_last_error() describes the last failure of
 * the calling thread.
 */
#ifndef BMFR_IO_H
#define BMFR_IO_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    BMFR_IO_OK = 0,
    BMFR_IO_ERR_ARGUMENT = -1,
    BMFR_IO_ERR_OPEN = -2,        /* file missing / unreadable / not writable */
    BMFR_IO_ERR_FORMAT = -3,      /* not an OpenEXR scanline file, truncated, corrupt chunk */
    BMFR_IO_ERR_UNSUPPORTED = -4, /* valid file this reader does not cover (tiled, deep, multi-part, PXR24/B44/DWA ...) */
    BMFR_IO_ERR_MISMATCH = -5     /* size or channel count differs from what the caller asked for */
};

/* Width, height (of the data window) and channel count of an OpenEXR file: lifts the reference's
 * "TODO detect IMAGE_SIZES automatically from the input files" (bmfr.cpp:37-40). */
int bmfr_io_exr_info(const char* path, int* width, int* height, int* channels);

/* read_image_file(), bmfr.cpp:145-165: the file must hold exactly three channels of width x height pixels
 * (else BMFR_IO_ERR_MISMATCH, the reference's "Can't open image file or it has wrong type"); HALF channels
 * are widened to fp32 (bmfr.cpp:159-161).  rgb receives width*height*3 tightly packed interleaved floats,
 * top row first.  Channels are delivered in R,G,B order when they are named so (as OpenImageIO does), X,Y,Z
 * likewise, otherwise in the file's (alphabetical) order.
 * Covered: single-part scanline files, compression NONE / RLE / ZIPS / ZIP / PIZ (the lossless ones), HALF and
 * FLOAT channels. */
int bmfr_io_read_exr_rgb(const char* path, int width, int height, float* rgb);

/* The dataset's camera_matrices.h (bmfr.cpp:46-47), parsed instead of compiled in: the initialisers of
 *   camera_matrices[frames][4][4]  -> matrices[frame*16 + row*4 + col]   (bound per frame at bmfr.cpp:440-442)
 *   pixel_offsets[frames][2]       -> offsets[frame*2 + i]               (bmfr.cpp:443-444)
 *   position_limit_squared, normal_limit_squared                         (bmfr.cpp:226-227)
 * At most max_frames entries are stored; n_matrices / n_offsets receive how many the file holds.  A limit the
 * file does not define is left untouched (and is not an error); missing arrays are BMFR_IO_ERR_FORMAT. */
int bmfr_io_parse_camera_header(const char* path, int max_frames, float* matrices, float* offsets, int* n_matrices,
                                int* n_offsets, float* position_limit_squared, float* normal_limit_squared);

/* The output images, bmfr.cpp:520-539: width x height pixels cropped out of rows of row_stride_floats floats
 * (the reference passes WORKSET_WIDTH*3), written as 8-bit RGB PNG: clamp to [0,1], *255, round to nearest
 * (NaN -> 0).  The values are the tone-mapped result of the taa kernel, no further transfer curve. */
int bmfr_io_write_png_rgb(const char* path, int width, int height, const float* rgb, size_t row_stride_floats);

/* Image-quality metrics of the paper's evaluation, for comparing output frames with ground truth:
 *   PSNR = 10 log10(peak^2 / mean squared error) over n values (+inf for identical inputs), NaNs are an error;
 *   SSIM = mean structural similarity (Wang et al. 2004: 11x11 Gaussian window, sigma 1.5, K1 = 0.01, K2 = 0.03,
 *          dynamic range `peak`, valid windows only), per channel of an interleaved width x height x 3 image, averaged.
 * tone_map applies the display transform of the accumulate_filtered_data kernel to linear radiance in place:
 * clamp(pow(max(0, v), 0.454545), 0, 1) (bmfr.cl:852-856), so that a linear ground-truth image can be compared with
 * the denoiser's tone-mapped output. */
int bmfr_io_psnr(const float* a, const float* b, size_t n, float peak, double* psnr_db);
int bmfr_io_ssim_rgb(const float* a, const float* b, int width, int height, float peak, double* ssim);
void bmfr_io_tone_map(float* rgb, size_t n);

const char* bmfr_io_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
