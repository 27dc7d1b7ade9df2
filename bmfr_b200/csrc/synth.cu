// synth-v1 generator: host twin + CUDA twin + C entry points (declared in include/bmfr_b200.h).
// Stands in for the dataset loader of the reference (/root/reference/opencl/bmfr.cpp:259-307) and
// for camera_matrices.h (bmfr.cpp:47,226-227,440-444), neither of which ships with the repository.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/bmfr_b200.h"
#include "bmfr_error.h"
#include "synth_core.h"

// Camera path.  Trigonometry in double on the host only; the kernels see rounded floats, so the
// host and device twins consume identical bits.
static void synth_camera_impl(int frame, int W, int H, int jitter, SynthCamera* cam, float* M, float* off) {
    const double f = (double)frame;
    const double o[3] = {-2.0 + 0.03 * f, 2.2 + 0.004 * f, -9.0 + 0.02 * f};
    const double yaw = 0.10 + 0.002 * f, pitch = -0.10;
    const double fw[3] = {sin(yaw) * cos(pitch), sin(pitch), cos(yaw) * cos(pitch)};
    const double rt[3] = {cos(yaw), 0.0, -sin(yaw)};
    // up = fwd x right ; down = -up
    const double up[3] = {fw[1] * rt[2] - fw[2] * rt[1], fw[2] * rt[0] - fw[0] * rt[2], fw[0] * rt[1] - fw[1] * rt[0]};
    const double thx = 0.7, thy = 0.7 * (double)H / (double)W;
    SynthCamera c;
    for (int k = 0; k < 3; ++k) {
        c.o[k] = (float)o[k];
        c.r[k] = (float)rt[k];
        c.d[k] = (float)(-up[k]);
        c.f[k] = (float)fw[k];
    }
    c.thx = (float)thx;
    c.thy = (float)thy;
    if (cam) *cam = c;
    if (M) {
        // clip_j = sum_i p_i * M[i][j], p_3 = 1 (bmfr.cl:343-347 picks columns .s048c/.s159d/.s37bf)
        double dn[3] = {-up[0], -up[1], -up[2]};
        double od_r = 0, od_d = 0, od_f = 0;
        for (int k = 0; k < 3; ++k) {
            od_r += (double)c.o[k] * rt[k];
            od_d += (double)c.o[k] * dn[k];
            od_f += (double)c.o[k] * fw[k];
        }
        for (int i = 0; i < 3; ++i) {
            M[i * 4 + 0] = (float)(rt[i] / thx);
            M[i * 4 + 1] = (float)(dn[i] / thy);
            M[i * 4 + 2] = (float)fw[i];
            M[i * 4 + 3] = (float)fw[i];
        }
        M[12] = (float)(-od_r / thx);
        M[13] = (float)(-od_d / thy);
        M[14] = (float)(-od_f - 0.1);
        M[15] = (float)(-od_f);
    }
    if (off) {
        off[0] = 0.5f;
        off[1] = 0.5f;
        if (jitter) {
            const uint32_t h = synth_hash(0xA511E9B3u ^ (uint32_t)frame);
            off[0] = 0.5f + 0.5f * ((float)(h & 0xFFFFu) * (1.f / 65536.f) - 0.5f);
            off[1] = 0.5f + 0.5f * ((float)(h >> 16) * (1.f / 65536.f) - 0.5f);
        }
    }
}

__global__ void synth_frame_kernel(SynthCamera cam, int W, int H, int y0, int y1, int frame, uint32_t seed,
                                   float* __restrict__ albedo, float* __restrict__ normal,
                                   float* __restrict__ position, float* __restrict__ color) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= y1) return;
    const SynthPixel p = synth_pixel(cam, W, H, x, y, frame, seed);
    const size_t i = ((size_t)(y - y0) * W + x) * 3;
    for (int k = 0; k < 3; ++k) {
        albedo[i + k] = p.albedo[k];
        normal[i + k] = p.normal[k];
        position[i + k] = p.position[k];
        color[i + k] = p.color[k];
    }
}

extern "C" {

void bmfr_synth_camera(int frame, int width, int height, int jitter, float cam_matrix[16], float pixel_offset[2]) {
    synth_camera_impl(frame, width, height, jitter, nullptr, cam_matrix, pixel_offset);
}

void bmfr_synth_limits(float* position_limit_squared, float* normal_limit_squared) {
    // plays the role of the two constants in the dataset's camera_matrices.h (bmfr.cpp:226-227)
    if (position_limit_squared) *position_limit_squared = 0.0225f;
    if (normal_limit_squared) *normal_limit_squared = 0.04f;
}

int bmfr_synth_frame_host(int width, int height, int y0, int y1, int frame, unsigned seed, float* albedo,
                          float* normal, float* position, float* noisy) {
    if (width <= 0 || height <= 0 || y0 < 0 || y1 > height || y0 > y1 || !albedo || !normal || !position || !noisy)
        return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_synth_frame_host: bad arguments");
    SynthCamera cam;
    synth_camera_impl(frame, width, height, 0, &cam, nullptr, nullptr);
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = y0; y < y1; ++y)
        for (int x = 0; x < width; ++x) {
            const SynthPixel p = synth_pixel(cam, width, height, x, y, frame, seed);
            const size_t i = ((size_t)(y - y0) * width + x) * 3;
            for (int k = 0; k < 3; ++k) {
                albedo[i + k] = p.albedo[k];
                normal[i + k] = p.normal[k];
                position[i + k] = p.position[k];
                noisy[i + k] = p.color[k];
            }
        }
    return BMFR_OK;
}

int bmfr_synth_frame_device(int width, int height, int y0, int y1, int frame, unsigned seed, float* d_albedo,
                            float* d_normal, float* d_position, float* d_noisy, void* stream) {
    if (width <= 0 || height <= 0 || y0 < 0 || y1 > height || y0 >= y1 || !d_albedo || !d_normal || !d_position ||
        !d_noisy)
        return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_synth_frame_device: bad arguments");
    SynthCamera cam;
    synth_camera_impl(frame, width, height, 0, &cam, nullptr, nullptr);
    dim3 block(32, 8), grid((width + 31) / 32, (y1 - y0 + 7) / 8);
    synth_frame_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(cam, width, height, y0, y1, frame, seed, d_albedo,
                                                                 d_normal, d_position, d_noisy);
    return bmfr_check_cuda(cudaGetLastError(), "synth_frame_kernel launch");
}

}  // extern "C"
