// Kernel parameter block and launcher prototypes shared by the pipeline and the kernels.
#pragma once
#include <cuda_runtime.h>

#include "../../include/bmfr_b200.h"

#define BMFR_FIT_THREADS 256  // LOCAL_SIZE, bmfr.cpp:116
#define BMFR_ROWS_PER_THREAD 4

// Strip contexts with connected neighbours: what one kernel of a frame (the reprojection, or the post pass) does for the
// halo exchange itself (option A: peer stores from inside the kernels).  Only the CTAs of the "zone" — those
// that cover rows within halo_rows of a strip edge with a neighbour — take part: they are the only ones whose gathers can
// reach halo rows and the only ones that own rows a neighbour mirrors.
//   prologue : one thread polls this context's flags until the neighbours' rows of the frames named below have arrived
//              (and the neighbours are done reading what this kernel's pushes overwrite);
//   body     : every store to a row of push_y0..push_y1 is repeated into the neighbour's halo (peer-mapped memory);
//   epilogue : the last zone CTA to finish raises the neighbours' flag to signal_value.
struct HaloK {
    unsigned int* flags;            // this context's flags: [0],[1] early from above / below, [4],[5] late, [2] a wait timed out
    unsigned int wait_early, wait_late;  // poll until early >= wait_early and late >= wait_late on every connected side (0: none)
    int side_on[2];                 // neighbour connected above / below
    int zone_y[2];                  // rows y < zone_y[0] or y >= zone_y[1] belong to the zone
    int push_y0[2], push_y1[2];     // image rows of this context stored a second time into the neighbour on that side
    float* peer_a[2];               // reprojection: the neighbour's accumulated noisy colour; post: its accumulated filtered colour
    float* peer_b[2];               // post: the neighbour's TAA result
    unsigned char* peer_c[2];       // reprojection: the neighbour's spp
    int peer_row0[2];               // first image row the neighbour's buffers hold
    unsigned int* peer_flag[2];     // the flag of the neighbour this kernel raises
    unsigned int signal_value;
    unsigned int* done_counter;     // zone CTAs of this launch that have finished (reset by the last one)
    unsigned int zone_ctas;
    int rows_top, rows_bot;         // CTA rows of the launch that belong to the zone at its upper / lower end: they are scheduled first
    unsigned long long timeout_ns;
    int active;                     // 0: no neighbour, nothing of the above happens
};

// Everything a kernel of one frame needs; passed by value as a __grid_constant__.
struct KParams {
    int W, H;            // IMAGE_WIDTH / IMAGE_HEIGHT (full image)
    int row0, row1;      // image rows held in the per-pixel buffers; buffer row = y - row0
    int frame;           // frame_number
    int off_x, off_y;    // BLOCK_OFFSETS[frame % 16], bmfr.cl:267-285
    int blocks_x;        // WORKSET_WITH_MARGINS_WIDTH / 32
    int blocks_y;        // WORKSET_WITH_MARGINS_HEIGHT / 32
    int by0, by1;        // block rows to process this frame (whole image: 0..blocks_y)
    int py0, py1;        // image rows the per-pixel stages (K3,K4) cover; K5 covers own_y0..own_y1
    int own_y0, own_y1;
    int k1_y0, k1_y1;    // image rows reproject_kernel covers: every row a block of by0..by1 reads
    int state2_row0, state2_row1;  // rows of accum / result that are valid on this strip (own rows + refreshed halo)
    float cam[16];       // prev_frame_camera_matrix
    float poff_x;        // pixel_offset.x
    float poff_y1;       // 1 - pixel_offset.y  (bmfr.cl:353-355)
    float blend_alpha, second_blend_alpha, taa_blend_alpha;
    float pos_limit, nrm_limit;
    const float* cur_normals;
    const float* prev_normals;
    const float* cur_positions;
    const float* prev_positions;
    const float* cur_noisy;       // this frame's 1-spp input (read-only)
    const float* prev_noisy_acc;  // previous_noisy
    float* cur_noisy_acc;         // what the reference stores back into current_noisy (bmfr.cl:481)
    const unsigned char* prev_spp;
    unsigned char* cur_spp;
    float2* prev_pixels;          // out_prev_frame_pixel / in_prev_frame_pixel
    unsigned char* accept;        // accept_bools
    float* tmp_data;              // STAGED only (fp16 elements when tmp_half)
    int tmp_half, reference_order;
    float* weights;
    float* mins_maxs;
    float* mins_inv;              // (min, 1/range or 1) per block and scaled feature, see scale_factor()
    const double* noise;          // [9][1024] add_random() increments of this frame
    const float* noise_f;         // the same tile rounded to fp32 (FUSED fit)
    double* noise_out;            // FUSED: the reproject kernel writes both tiles itself
    float* noise_f_out;
    double noise_amount;          // NOISE_AMOUNT (a double literal in the reference, bmfr.cpp:58)
    const float* albedo;
    float* filtered;              // STAGED only
    const float* accum_prev;      // accumulated_prev_frame
    float* accum_cur;             // accumulated_frame
    float* tone_mapped;           // STAGED only
    const float* result_prev;     // prev_frame (taa)
    float* result_cur;            // result_frame
    float* user_out;              // optional copy of result rows for the caller
    int* oob_flag;                // set when a gather needed a row outside [row0,row1)
    int* block_counter;           // FUSED fit: dynamic block schedule, reset by the reprojection (STAGED: noise-tile kernel) of the frame
    float* tri;                   // FUSED fit: level-1 triangles between the two levels of the TSQR, blocks x 4 x 136 floats
    // profile = 2: per kernel (reproject, fit, post) the earliest CTA start and the latest CTA end of this frame's launch,
    // globaltimer ns, both kept with atomicMin (the end as ~t); stamps_next = the next frame's slot, re-armed by the reprojection
    unsigned long long* stamps;
    unsigned long long* stamps_next;
    HaloK halo_r, halo_p;         // halo exchange duties of the reprojection / the post pass (FUSED strips)
    int feature_set;              // bmfr_feature_set: which instantiation of the FUSED fit / post pass runs
    int n_features;               // its feature count F (the noise tile has F - 1 columns, BUFFER_COUNT = F + 3)
    int n_scaled;                 // its scaled features
    int fit_method;               // host side only: BMFR_FIT_GRAM / BMFR_FIT_TSQR (which FUSED fit kernel launch_fit_qr starts)
    int plain_launch;             // host side only: launch the FUSED kernels without programmatic stream serialization
    int hist_stride;              // floats per pixel of accum_* / result_*: 3 (bmfr.cl:224-241), or 4 = padded to 16 bytes (whole-image
                                  // FUSED contexts on the TMA post pass, see load_hist() in bmfr_post.cu)
};


void bmfr_host_block_offset(int frame, int* ox, int* oy);
cudaError_t launch_noise_tile(double* d_noise, float* d_noise_f, int* block_counter, double noise_amount, int frame, cudaStream_t st);
cudaError_t launch_k1(const KParams& P, cudaStream_t st);
cudaError_t launch_k2(const KParams& P, cudaStream_t st);
cudaError_t launch_k2_reference_order(const KParams& P, bool half, cudaStream_t st);
cudaError_t launch_k3_reference_order(const KParams& P, cudaStream_t st);
cudaError_t launch_k3(const KParams& P, cudaStream_t st);
cudaError_t launch_k4(const KParams& P, cudaStream_t st);
cudaError_t launch_k5(const KParams& P, cudaStream_t st);
cudaError_t launch_reproject(const KParams& P, cudaStream_t st);
cudaError_t launch_fit_qr(const KParams& P, cudaStream_t st);
cudaError_t launch_post(const KParams& P, cudaStream_t st);
bool post_uses_tma(int width, int rows);
