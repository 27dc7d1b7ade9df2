// sm_100a kernels of the BMFR per-frame path and their launchers.
//   STAGED : the reference's five kernels (bmfr.cl:290,490,703,761,860), one launch each
//   FUSED  : bmfr_fit.cu (reproject + fit) and bmfr_post.cu (weighted sum + accumulation + taa)
// Compiled with --fmad=false; see bmfr_device.cuh for the arithmetic convention.
#include <cuda_fp16.h>

#include "bmfr_kernels.h"

#include "bmfr_device.cuh"

// --------------------------------------------------------------------------------------------
// add_random() increments of one frame (bmfr.cl:173-182).  The seed does not contain the block
// index, so all blocks of a frame share one 9x1024 tile; it is computed once per frame instead of
// once per element per block.  Kept in fp64: NOISE_AMOUNT is a double literal in the reference
// (bmfr.cpp:58), which makes `value + NOISE_AMOUNT * 2.f * (random - 0.5f)` an fp64 expression.
// --------------------------------------------------------------------------------------------
__global__ void noise_tile_kernel(double* __restrict__ noise, float* __restrict__ noise_f, int* __restrict__ block_counter,
                                  double noise_amount, int frame) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *block_counter = 0;  // the fit kernel of this frame draws its blocks from it
    if (i >= (BMFR_FEATURES - 1) * BMFR_BLOCK_PIXELS) return;
    // id + sub_vector*256 + feature_buffer*1024 + frame*13*1024 with feature_buffer = 1 + i/1024
    const int seed = i + BMFR_BLOCK_PIXELS + frame * BMFR_BUFFER_COUNT * BMFR_BLOCK_PIXELS;
    const float r = bmfr_random((unsigned int)seed) - 0.5f;
    const double d = (noise_amount * 2.0) * (double)r;
    noise[i] = d;
    noise_f[i] = (float)d;
}

// --------------------------------------------------------------------------------------------
// STAGED K1: one thread per work-item of the margin-extended domain (bmfr.cl:310-484).
// --------------------------------------------------------------------------------------------
// HALF: USE_HALF_PRECISION_IN_TMP_DATA = 1 — clamp to the fp16 range, store rounded to fp16 (bmfr.cl:471-473)
template <bool STRIP, bool HALF>
__global__ void __launch_bounds__(256) k1_accumulate_noisy_kernel(const __grid_constant__ KParams P) {
    const int gx = blockIdx.x * 32 + threadIdx.x;
    const int gy = P.by0 * 32 + blockIdx.y * 8 + threadIdx.y;
    const int ux = gx - 16 + P.off_x, uy = gy - 16 + P.off_y;  // pixel_without_mirror
    const int x = mirror_index(ux, P.W), y = mirror_index(uy, P.H);
    if (y < P.row0 || y >= P.row1) {  // only possible on a strip whose halo is too small
        *P.oob_flag = 1;
        return;
    }
    const K1Pixel r = k1_pixel<STRIP>(P, x, y);
    float f[BMFR_BUFFER_COUNT];
    k1_features(r, f);
    const int bx = gx >> 5, by = gy >> 5;
    const size_t t0 = ((size_t)((by - P.by0) * P.blocks_x + bx) * BMFR_BUFFER_COUNT) * BMFR_BLOCK_PIXELS + (gy & 31) * 32 + (gx & 31);
#pragma unroll
    for (int i = 0; i < BMFR_BUFFER_COUNT; ++i) {
        if (HALF)
            reinterpret_cast<__half*>(P.tmp_data)[t0 + (size_t)i * BMFR_BLOCK_PIXELS] =
                __float2half_rn(fmaxf(fminf(f[i], 65504.f), -65504.f));
        else
            P.tmp_data[t0 + (size_t)i * BMFR_BLOCK_PIXELS] = f[i];
    }
    if (ux >= 0 && ux < P.W && uy >= 0 && uy < P.H) {  // bmfr.cl:478-484
        const size_t lp = pix_index(P, x, y);
        store_f3(P.cur_noisy_acc, lp, r.new_color);
        P.cur_spp[lp] = r.spp;  // the reference also lets mirrored work-items store it (same value)
        P.prev_pixels[lp] = make_float2(r.prev_x, r.prev_y);
        P.accept[lp] = r.accept;
    }
}

#ifndef BMFR_FIT_MIN_BLOCKS
#define BMFR_FIT_MIN_BLOCKS 2
#endif
// STAGED K2: one CTA per block, reads K1's block-planar tmp_data (bmfr.cl:490-700).
__global__ void __launch_bounds__(BMFR_FIT_THREADS, BMFR_FIT_MIN_BLOCKS) k2_fitter_kernel(const __grid_constant__ KParams P) {
    __shared__ __align__(16) FitShared sh;
    const int bx = blockIdx.x, by = P.by0 + blockIdx.y;
    const int group = by * P.blocks_x + bx;
    const float* t = P.tmp_data + ((size_t)((by - P.by0) * P.blocks_x + bx) * BMFR_BUFFER_COUNT) * BMFR_BLOCK_PIXELS;
    float a[BMFR_ROWS_PER_THREAD][BMFR_BUFFER_COUNT];
#pragma unroll
    for (int s = 0; s < BMFR_ROWS_PER_THREAD; ++s)
#pragma unroll
        for (int c = 0; c < BMFR_BUFFER_COUNT; ++c)
            a[s][c] = t[(size_t)c * BMFR_BLOCK_PIXELS + threadIdx.x + BMFR_FIT_THREADS * s];
    block_fit(a, sh, P.noise, P.weights, P.mins_maxs, P.mins_inv, group);
}

// --------------------------------------------------------------------------------------------
// STAGED K3 / K4 / K5: one thread per image pixel (bmfr.cl:703-758, 761-857, 860-974).
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k3_weighted_sum_kernel(const __grid_constant__ KParams P) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = P.py0 + blockIdx.y * 8 + threadIdx.y;
    if (x >= P.W || y >= P.py1) return;
    const size_t lp = pix_index(P, x, y);
    const int g = k3_group(P, x, y);
    const f3 c = k3_pixel(load_f3(P.cur_normals, lp), load_f3(P.cur_positions, lp),
                          P.weights + (size_t)g * BMFR_FEATURES * 3, P.mins_inv + (size_t)g * BMFR_FEATURES_SCALED * 2);
    store_f3(P.filtered, lp, c);
}

__global__ void __launch_bounds__(256) k4_accumulate_filtered_kernel(const __grid_constant__ KParams P) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = P.py0 + blockIdx.y * 8 + threadIdx.y;
    if (x >= P.W || y >= P.py1) return;
    const size_t lp = pix_index(P, x, y);
    const float2 pp = P.prev_pixels[lp];
    f3 accum, tone;
    k4_pixel(P, lp, load_f3(P.filtered, lp), pp.x, pp.y, P.accept[lp], accum, tone);
    store_f3(P.accum_cur, lp, accum);
    store_f3(P.tone_mapped, lp, tone);
}

__global__ void __launch_bounds__(256) k5_taa_kernel(const __grid_constant__ KParams P) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = P.own_y0 + blockIdx.y * 8 + threadIdx.y;
    if (x >= P.W || y >= P.own_y1) return;
    const size_t lp = pix_index(P, x, y);
    const f3 my_new = load_f3(P.tone_mapped, lp);
    const float2 pp = P.prev_pixels[lp];
    const int pix = __float2int_rd(pp.x), piy = __float2int_rd(pp.y);
    f3 out;
    if (P.frame == 0 || pix < -1 || piy < -1 || pix >= P.W || piy >= P.H) {  // bmfr.cl:884-890
        out = my_new;
    } else {
        TaaBox box;
        taa_box_init(box);
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int sx = x + dx, sy = y + dy;
                if (sx >= 0 && sy >= 0 && sx < P.W && sy < P.H) {
                    const f3 s = (dx == 0 && dy == 0) ? my_new : load_f3(P.tone_mapped, pix_index(P, sx, sy));
                    taa_box_add(box, s, dx == 0 || dy == 0);
                }
            }
        out = taa_resolve(P, my_new, box, pp.x, pp.y, pix, piy);
    }
    store_f3(P.result_cur, lp, out);
    if (P.user_out) store_f3(P.user_out, lp, out);
}

// --------------------------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------------------------
static const int h_block_offsets[BMFR_BLOCK_OFFSETS_COUNT][2] = {  // BLOCK_OFFSETS, bmfr.cl:268-285
    {-14, -14}, {4, -6}, {-8, 14}, {8, 0},   {-10, -8}, {2, 12},  {12, -12}, {-10, 0},
    {12, 14},   {-8, -16}, {6, 6}, {-2, -2}, {6, -14},  {-16, 12}, {14, -4}, {-6, 4}};

void bmfr_host_block_offset(int frame, int* ox, int* oy) {
    const int i = ((frame % BMFR_BLOCK_OFFSETS_COUNT) + BMFR_BLOCK_OFFSETS_COUNT) % BMFR_BLOCK_OFFSETS_COUNT;
    *ox = h_block_offsets[i][0];
    *oy = h_block_offsets[i][1];
}

cudaError_t launch_noise_tile(double* d_noise, float* d_noise_f, int* block_counter, double noise_amount, int frame, cudaStream_t st) {
    const int n = (BMFR_FEATURES - 1) * BMFR_BLOCK_PIXELS;
    noise_tile_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_noise, d_noise_f, block_counter, noise_amount, frame);
    return cudaGetLastError();
}

// a context that holds only a band of rows (strip + halo) checks its gathers against the band
static bool is_strip(const KParams& P) { return P.row0 != 0 || P.row1 != P.H; }
static dim3 pixel_grid(const KParams& P, int y0, int y1) { return dim3((P.W + 31) / 32, (y1 - y0 + 7) / 8); }

cudaError_t launch_k1(const KParams& P, cudaStream_t st) {
    dim3 grid(P.blocks_x, (P.by1 - P.by0) * 4), block(32, 8);
    if (P.tmp_half) {
        if (is_strip(P)) k1_accumulate_noisy_kernel<true, true><<<grid, block, 0, st>>>(P);
        else k1_accumulate_noisy_kernel<false, true><<<grid, block, 0, st>>>(P);
    } else {
        if (is_strip(P)) k1_accumulate_noisy_kernel<true, false><<<grid, block, 0, st>>>(P);
        else k1_accumulate_noisy_kernel<false, false><<<grid, block, 0, st>>>(P);
    }
    return cudaGetLastError();
}
cudaError_t launch_k2(const KParams& P, cudaStream_t st) {
    dim3 grid(P.blocks_x, P.by1 - P.by0);
    k2_fitter_kernel<<<grid, BMFR_FIT_THREADS, 0, st>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_k3(const KParams& P, cudaStream_t st) {
    k3_weighted_sum_kernel<<<pixel_grid(P, P.py0, P.py1), dim3(32, 8), 0, st>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_k4(const KParams& P, cudaStream_t st) {
    k4_accumulate_filtered_kernel<<<pixel_grid(P, P.py0, P.py1), dim3(32, 8), 0, st>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_k5(const KParams& P, cudaStream_t st) {
    k5_taa_kernel<<<pixel_grid(P, P.own_y0, P.own_y1), dim3(32, 8), 0, st>>>(P);
    return cudaGetLastError();
}
