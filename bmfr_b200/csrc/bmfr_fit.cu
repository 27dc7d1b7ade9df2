// FUSED mode, first half of the frame: accumulate_noisy_data + fitter (bmfr.cl:290-485, 490-700)
// as two sm_100a kernels sized for what each phase is bound by:
//
//   reproject_kernel  : K1 for every image pixel, four pixels per thread with the next pixel's world
//                       position fetched one pixel ahead (gather-latency bound).  Writes the four
//                       per-pixel outputs of bmfr.cl:478-484 and, in its first CTAs, the frame's noise
//                       tile.  The block-planar tmp_data of the reference is never written.
//   fit_qr_kernel     : persistent four-warp CTAs (four per SM) walking over the 32x32 blocks.  The
//                       warps rebuild the block's 1024x13 matrix in registers from the per-pixel
//                       buffers (mirrored margins included, bmfr.cl:314-316; the three 32x32-pixel
//                       input tiles of the next block arrive by TMA while the current block is
//                       factored), block min/max + scaling + noise (bmfr.cl:511-542, 623-627), then
//                       each factors its own 256 rows (level 1 of a TSQR) without a block barrier.
//                       After its last block the CTA factors the four stacked triangles of each of
//                       its blocks (level 2) and back-substitutes (bmfr.cl:659-699).
//
// Compiled with --fmad=false (K1 is bit-exact against the oracle); the fit writes fmaf() explicitly.
#include "bmfr_kernels.h"

#include "bmfr_device.cuh"
#include "bmfr_tma.cuh"

// --------------------------------------------------------------------------------------------
// K1 per image pixel.  Mirrored margin work-items of the reference recompute an in-image pixel
// (bmfr.cl:314-325) and store nothing per pixel (bmfr.cl:478), so the image pixels are the whole
// job; the fit re-derives margin rows from these outputs.
// --------------------------------------------------------------------------------------------
// CTAs per SM and CTA shape of the whole-image instantiation.  Measured at 1080p (profiles/r02_u_*): 4 CTAs x (32 x 8
// threads) 52.1 us; 5 CTAs (48 registers, no spills) 48.4 us; 5 CTAs x (64 x 4 threads) 48.0 us; 6 CTAs (40 registers,
// spills) 65.5 us.  The strip instantiation keeps 4 x (32 x 8): its row staging is laid out for 32 x 32 tiles and it
// spills at 48 registers.
#ifndef BMFR_REPROJECT_MIN_BLOCKS
#define BMFR_REPROJECT_MIN_BLOCKS 5
#endif
#ifndef BMFR_REPROJECT_BX
#define BMFR_REPROJECT_BX 64  // CTA = BX x (256 / BX) threads
#endif
#ifndef BMFR_REPROJECT_STRIP_MIN_BLOCKS
#define BMFR_REPROJECT_STRIP_MIN_BLOCKS 4
#endif
#define BMFR_REPROJECT_STRIP_BX 32
// 1: a thread walks down four vertically adjacent pixels and hands the lower tap row of one to the next (K1Carry; 25 %
// fewer tap loads).  Measured at 1080p (profiles/r02_m_*): 69.8 us against 52 us — the eight warps of a CTA then work on
// rows four apart instead of on eight consecutive rows (the adjacent-row mapping alone: 58.1 us), which costs more L1
// locality than the shared rows save, and the carried row spills.  Off.
#ifndef BMFR_REPROJECT_CARRY
#define BMFR_REPROJECT_CARRY 0
#endif
#ifndef BMFR_REPROJECT_PIXELS
#define BMFR_REPROJECT_PIXELS 4  // pixels per thread (rows BY apart); the next pixel's position is fetched one pixel ahead
#endif
// Layout of the fp32 copy of the noise tile (read only by fit_qr_kernel): feature c (0..8 = columns
// 1..9), pixel (x, y) of the block.  The fit's thread (warp = y / 8, lane = x) owns rows y = 8 warp .. +7
// and reads them as two float4 per feature, consecutive lanes 16 bytes apart.
__device__ __forceinline__ int noise_f_index(int c, int y, int x, int noisy_columns) {
    return (((((y >> 3) * noisy_columns + c) * 2 + ((y & 7) >> 2)) * 32 + x) << 2) + (y & 3);
}

// The first CTAs of the reprojection also produce this frame's add_random() tile (bmfr.cl:173-182; one 9x1024 tile per
// frame shared by all blocks, fp64 like the reference's double literal plus its fp32 rounding for the fit) and reset the
// fit's block counter: the fit starts only after the reprojection has completed.
__device__ __forceinline__ void reproject_noise_tile(const KParams& P, int tid) {
    const int cta = blockIdx.y * gridDim.x + blockIdx.x, ncta = gridDim.x * gridDim.y;
    const int workers = ncta < 36 ? ncta : 36;
    if (cta < workers) {
        const int n = (P.n_features - 1) * BMFR_BLOCK_PIXELS;  // columns 1 .. F-1 get noise; BUFFER_COUNT = F + 3 (bmfr.cl:179-181)
        for (int i = cta * 256 + tid; i < n; i += workers * 256) {
            const int seed = i + BMFR_BLOCK_PIXELS + P.frame * (P.n_features + 3) * BMFR_BLOCK_PIXELS;
            const double d = (P.noise_amount * 2.0) * (double)(bmfr_random((unsigned int)seed) - 0.5f);
            P.noise_out[i] = d;
            P.noise_f_out[noise_f_index(i / BMFR_BLOCK_PIXELS, (i % BMFR_BLOCK_PIXELS) / 32, i % 32, P.n_features - 1)] = (float)d;
        }
        if (cta == 0 && tid == 0) *P.block_counter = 0;
    }
}

// this frame's normals / world positions (the fit and the post pass read them again) and its 1-spp colour (read once)
__device__ __forceinline__ f3 load_cur(const float* __restrict__ b, unsigned int i) {
#if BMFR_L2_HINTS
    return load_f3_hint(b, i, BMFR_L2_KEEP);
#else
    return load_f3_stream(b, i);
#endif
}
__device__ __forceinline__ f3 load_once(const float* __restrict__ b, unsigned int i) {
#if BMFR_L2_HINTS >= 2
    return load_f3_hint(b, i, BMFR_L2_ONCE);
#else
    return load_f3_stream(b, i);
#endif
}

// Zone CTAs of a strip stage the rows a neighbour mirrors in shared memory and send them as 16-byte peer stores, one
// full 384-byte row segment per 24 lanes: single 4-byte stores at a 12-byte stride make poor NVLink packets, and at 8K
// over eight GPUs a rank pushes ~12 MB per frame from this kernel.  (Whole-image instantiations carry no staging.)
template <bool STRIP, bool WITH_RGB>
struct ReprojectPushStage {
    float rgb[WITH_RGB ? 32 : 1][96];
    unsigned char spp[32][32];
};
template <bool WITH_RGB>
struct ReprojectPushStage<false, WITH_RGB> {
    float rgb[1][96];  // (never touched)
    unsigned char spp[1][32];
};

// All threads of a zone CTA, after a barrier: rows [cta_y0, cta_y0 + 32) x columns [x0, x0 + 32) of the stage -> the
// neighbours that mirror them.  Needs W % 32 == 0 (full tiles, 16-byte aligned row segments).
__device__ __forceinline__ void reproject_push_rows(const KParams& P, const float (*rgb)[96], const unsigned char (*spp)[32], int x0, int cta_y0, int tid) {
    const HaloK& h = P.halo_r;
#ifdef BMFR_DEBUG_NO_PUSH
    return;
#endif
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        if (!h.side_on[s]) continue;
        const int ya = max(h.push_y0[s], cta_y0), yb = min(h.push_y1[s], cta_y0 + 32);
        const int items = (yb - ya) * 26;  // per row: 24 x 16 B of colour, 2 x 16 B of sample counts
        for (int i = tid; i < items; i += 256) {
            const int y = ya + i / 26, k = i % 26, r = y - cta_y0;
            const long long pix = (long long)(y - h.peer_row0[s]) * P.W + x0;
            if (k < 24) reinterpret_cast<float4*>(h.peer_a[s] + pix * 3)[k] = reinterpret_cast<const float4*>(&rgb[r][0])[k];
            else reinterpret_cast<uint4*>(h.peer_c[s] + pix)[k - 24] = reinterpret_cast<const uint4*>(&spp[r][0])[k - 24];
        }
    }
}

// The four per-pixel outputs of bmfr.cl:478-484; in the zone of a strip the accumulated colour and the sample count of a
// row a neighbour mirrors go to its halo as well (peer memory over NVLink).
template <bool STRIP>
__device__ __forceinline__ void reproject_store(const KParams& P, int x, int y, const K1Pixel& r, bool zone, float (*st_rgb)[96],
                                                unsigned char (*st_spp)[32], int tx, int row_in_cta) {
    const unsigned int lp = pix_index(P, x, y);
#if BMFR_L2_HINTS
    store_f3_hint(P.cur_noisy_acc, lp, r.new_color, BMFR_L2_KEEP);
#else
    store_f3(P.cur_noisy_acc, lp, r.new_color);
#endif
    P.cur_spp[lp] = r.spp;
#if BMFR_L2_HINTS
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(P.prev_pixels + lp), "f"(r.prev_x), "f"(r.prev_y), "l"(BMFR_L2_KEEP) : "memory");
#else
    P.prev_pixels[lp] = make_float2(r.prev_x, r.prev_y);
#endif
    P.accept[lp] = r.accept;
    if constexpr (STRIP) {
        if (!zone) return;
        if ((P.W & 31) == 0) {  // staged, sent row by row after the CTA's last pixel (reproject_push_rows)
            st_rgb[row_in_cta][3 * tx] = r.new_color.x; st_rgb[row_in_cta][3 * tx + 1] = r.new_color.y; st_rgb[row_in_cta][3 * tx + 2] = r.new_color.z;
            st_spp[row_in_cta][tx] = r.spp;
            return;
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const long long pi = halo_peer_index(P.halo_r, P, s, x, y);
            if (pi >= 0) {
                float* pa = P.halo_r.peer_a[s] + pi * 3;
                pa[0] = r.new_color.x; pa[1] = r.new_color.y; pa[2] = r.new_color.z;
                P.halo_r.peer_c[s][pi] = r.spp;
            }
        }
    }
}

template <bool STRIP>
__global__ void __launch_bounds__(256, STRIP ? BMFR_REPROJECT_STRIP_MIN_BLOCKS : BMFR_REPROJECT_MIN_BLOCKS) reproject_kernel(const __grid_constant__ KParams P) {
    constexpr int BX = STRIP ? BMFR_REPROJECT_STRIP_BX : BMFR_REPROJECT_BX, BY = 256 / BX, CTA_ROWS = BY * BMFR_REPROJECT_PIXELS;
    static_assert(!STRIP || (BX == 32 && CTA_ROWS == 32), "the strip staging (ReprojectPushStage) holds 32 x 32 pixels");
    // Everything below reads the caller's inputs.  Their producer may be the kernel right before this one on the
    // context's stream — then it is this grid's programmatic-launch primary, and if it triggers its dependents early
    // its writes are only guaranteed visible after the wait.  So the wait comes first; what the programmatic launch
    // still buys is that this grid's CTAs are resident when the previous frame's post pass retires.
    const int x = blockIdx.x * BX + threadIdx.x;
    const int cta_y0 = P.k1_y0 + (STRIP ? halo_row_order(P.halo_r, blockIdx.y, gridDim.y) : sweep_row(P, blockIdx.y, gridDim.y)) * CTA_ROWS;
    // strips: a CTA near a strip edge waits for the neighbours' rows of the previous frame before it gathers from them; its
    // first look at the flags is in flight across the wait for the predecessor
    const bool zone = STRIP && halo_in_zone(P.halo_r, cta_y0, cta_y0 + CTA_ROWS);
    const HaloPeek peek = halo_peek(P.halo_r, zone);
    pdl_wait();
    // after the wait, so that "everything before this grid is complete" is transitive: the fit requests its first
    // normals / positions tiles (the caller's inputs) before its own wait
    pdl_trigger();  // the fit's CTAs may take SM slots as this grid drains
    stamp_begin(P, 0);
    if (P.stamps_next != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.y == 0 && threadIdx.x < 6) P.stamps_next[threadIdx.x] = ~0ull;
    reproject_noise_tile(P, threadIdx.y * BX + threadIdx.x);
    __shared__ __align__(16) ReprojectPushStage<STRIP, true> push_stage;
    // thread (x, ty) takes rows ty, ty + BY, ty + 2 BY, ty + 3 BY of the CTA's rows (BMFR_REPROJECT_CARRY: rows 4 ty .. 4 ty + 3)
    const int ystep = BMFR_REPROJECT_CARRY ? 1 : BY;
    const int ybase = cta_y0 + (BMFR_REPROJECT_CARRY ? BMFR_REPROJECT_PIXELS : 1) * threadIdx.y;
    int ylo = P.k1_y0, yhi = P.k1_y1;
    if (STRIP) {  // rows outside the strip + halo cannot be reprojected here: flag and skip them
        if (ylo < P.row0 || yhi > P.row1) *P.oob_flag = 1;
        ylo = max(ylo, P.row0);
        yhi = min(yhi, P.row1);
    }
    // software pipeline over the thread's pixels: position of pixel k+1 in flight while pixel k runs its
    // reprojection -> tap gather chain
    f3 wp_next = make_f3(0.f, 0.f, 0.f);
    if (x < P.W && ybase >= ylo && ybase < yhi) wp_next = load_cur(P.cur_positions, pix_index(P, x, ybase));
    if (zone) halo_poll(P.halo_r, peek);  // (this frame's own inputs never wait for a neighbour)
    if (x < P.W) {
        K1Carry carry;
        carry.ry = -1;
#pragma unroll 1
        for (int k = 0; k < BMFR_REPROJECT_PIXELS; ++k) {
            const int y = ybase + k * ystep;
            if (y >= yhi) break;
            const f3 wp = wp_next;
            const int yn = y + ystep;
            if (k + 1 < BMFR_REPROJECT_PIXELS && yn >= ylo && yn < yhi) wp_next = load_cur(P.cur_positions, pix_index(P, x, yn));
            if (y < ylo) continue;
            const unsigned int lp = pix_index(P, x, y);
            const K1Pixel r = k1_pixel_core<STRIP>(P, x, y, wp, load_cur(P.cur_normals, lp), load_once(P.cur_noisy, lp),
                                                   BMFR_REPROJECT_CARRY ? &carry : nullptr);
            reproject_store<STRIP>(P, x, y, r, zone, push_stage.rgb, push_stage.spp, threadIdx.x, y - cta_y0);
        }
    }
    if constexpr (STRIP) {
        if (zone) {
            if ((P.W & 31) == 0) {
                __syncthreads();
                reproject_push_rows(P, push_stage.rgb, push_stage.spp, blockIdx.x * BX, cta_y0, threadIdx.y * BX + threadIdx.x);
            }
            halo_finish(P.halo_r, halo_cta_pushes(P.halo_r, cta_y0, cta_y0 + 32));
        }
    }
    stamp_end(P, 0);
}

// --------------------------------------------------------------------------------------------
// reproject_tma_kernel: the same per-pixel arithmetic (k1_pixel_core), but a CTA's 32x32 pixels of the three
// current-frame inputs (world position, normal, noisy colour) arrive as three bulk tensor copies.  That takes the first
// of the two dependent memory round trips of a pixel (position -> reprojection -> 40 tap loads) off the critical path of
// every thread — the taps of a thread's first pixel are issued as soon as the tiles have landed — and leaves the load /
// store unit to the gathers.  Used when the tensor maps can be built (W % 4 == 0, at least 32 rows).
// --------------------------------------------------------------------------------------------
#define RP_TILE_W 96
struct ReprojectShared {
    float pos[32][RP_TILE_W];
    float nrm[32][RP_TILE_W];
    float col[32][RP_TILE_W];
    unsigned long long bar;
};
struct ReprojectMaps {
    CUtensorMap positions, normals, noisy;
};

template <bool STRIP>
__global__ void __launch_bounds__(256, BMFR_REPROJECT_STRIP_MIN_BLOCKS) reproject_tma_kernel(const __grid_constant__ KParams P,
                                                                                       const __grid_constant__ ReprojectMaps M) {
    __shared__ __align__(128) ReprojectShared sh;
    __shared__ __align__(16) ReprojectPushStage<STRIP, false> push_stage;  // the colour rows are staged over the noisy-colour tile
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x0 = blockIdx.x * 32, y0 = P.k1_y0 + (STRIP ? halo_row_order(P.halo_r, blockIdx.y, gridDim.y) : (int)blockIdx.y) * 32;
    if (threadIdx.x == 0) {
        mbar_init(&sh.bar, 1);
        mbar_fence_init();
    }
    pdl_wait();     // the caller's inputs are complete (see reproject_kernel)
    pdl_trigger();
    stamp_begin(P, 0);
    if (P.stamps_next != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 6) P.stamps_next[threadIdx.x] = ~0ull;
    if (threadIdx.x == 0) {  // rows below the image / the strip arrive as zeros and are never used
        mbar_expect_tx(&sh.bar, 3 * 32 * RP_TILE_W * 4);
        tma_load_tile(&sh.pos[0][0], &M.positions, x0 * 3, y0 - P.row0, &sh.bar);
        tma_load_tile(&sh.nrm[0][0], &M.normals, x0 * 3, y0 - P.row0, &sh.bar);
        tma_load_tile(&sh.col[0][0], &M.noisy, x0 * 3, y0 - P.row0, &sh.bar);
    }
    reproject_noise_tile(P, threadIdx.x);
    __syncthreads();  // the barrier's initialisation is visible
    const bool zone = STRIP && halo_in_zone(P.halo_r, y0, y0 + 32);
    if (zone) halo_poll(P.halo_r, halo_peek(P.halo_r, zone));
    int ylo = P.k1_y0, yhi = P.k1_y1;
    if (STRIP) {  // rows outside the strip + halo cannot be reprojected here: flag and skip them
        if (ylo < P.row0 || yhi > P.row1) *P.oob_flag = 1;
        ylo = max(ylo, P.row0);
        yhi = min(yhi, P.row1);
    }
    mbar_wait_hot(&sh.bar, 0);
    const int x = x0 + tx;
    if (x < P.W) {
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            const int r = ty + 8 * k, y = y0 + r;
            if (y >= yhi) break;
            if (y < ylo) continue;
            const f3 wp = make_f3(sh.pos[r][3 * tx], sh.pos[r][3 * tx + 1], sh.pos[r][3 * tx + 2]);
            const f3 n = make_f3(sh.nrm[r][3 * tx], sh.nrm[r][3 * tx + 1], sh.nrm[r][3 * tx + 2]);
            const f3 cur = make_f3(sh.col[r][3 * tx], sh.col[r][3 * tx + 1], sh.col[r][3 * tx + 2]);
            reproject_store<STRIP>(P, x, y, k1_pixel_core<STRIP>(P, x, y, wp, n, cur), zone, sh.col, push_stage.spp, tx, r);
        }
    }
    if constexpr (STRIP) {
        if (zone) {
            if ((P.W & 31) == 0) {
                __syncthreads();
                reproject_push_rows(P, sh.col, push_stage.spp, x0, y0, threadIdx.x);
            }
            halo_finish(P.halo_r, halo_cta_pushes(P.halo_r, y0, y0 + 32));
        }
    }
    stamp_end(P, 0);
}

// --------------------------------------------------------------------------------------------
// warp-level helpers
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_min(float v) {
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float warp_max(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
// three-input minimum / maximum (FMNMX3, sm_100)
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float rcp_approx(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float v) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// Sum over the 32 lanes of N (<= 16) per-lane values with the result of every column in every lane,
// using shuffles only.  Stage `bit` (16, 8, 4, 2, 1) folds lane bit `bit`: while a lane still carries
// more than one value the stage is a reduce-scatter step (the lane keeps one half of its values, sends
// the other half to its partner and adds what it receives), afterwards a plain butterfly step.  One
// indexed shuffle per column then broadcasts the totals.  16+N shuffles for N > 8 (9+N, 6+N for the
// narrow reflectors) instead of 5N for a butterfly on every value — and no shared-memory round trip
// and no warp barrier on the critical path of a reflector.
//   own : the total of the column this lane ended up holding, column index lane / (32 / W0).
template <int N>
struct AllSum {
    static constexpr int W0 = (N > 8) ? 16 : (N > 4) ? 8 : 4;
    static constexpr int LANES_PER_COLUMN = 32 / W0;
};
template <int N>
__device__ __forceinline__ void warp_allsum(float (&v)[16], int lane, float& own) {
    constexpr int W0 = AllSum<N>::W0;
    float w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = (i < N) ? v[i] : 0.f;
    int width = W0;
#pragma unroll
    for (int bit = 16; bit >= 1; bit >>= 1) {
        if (width > 1) {
            const int half = width / 2;
            const bool hi = (lane & bit) != 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < half) {
                    const float send = hi ? w[i] : w[i + half];
                    const float keep = hi ? w[i + half] : w[i];
                    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
                }
            }
            width = half;
        } else {
            w[0] += __shfl_xor_sync(0xffffffffu, w[0], bit);
        }
    }
    own = w[0];
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = __shfl_sync(0xffffffffu, w[0], j * AllSum<N>::LANES_PER_COLUMN);
}

// One reflector of level 1 against a virtual zero pivot row (see bmfr_device.cuh "The fit"): with
// S_j = a_k.a_j,   R_kj = S_j / sqrt(S_k),   a_j -= a_k * S_j / S_k   (j > k).
// a2[h][c-1] holds rows (2h, 2h+1) of column c (column 0 is the constant 1 and lives in no register),
// so the products and the eliminations are FFMA2; srow receives S_k..S_12, the unnormalised row k of R.
template <int PAIRS, int K>
__device__ __forceinline__ void qr_step2(float2 (&a2)[PAIRS][BMFR_BUFFER_COUNT - 1], float* __restrict__ srow, int lane) {
    constexpr int N = BMFR_BUFFER_COUNT - K;
    float cj[16];
    if (K == 0) {  // a_0 = 1: the products are plain column sums, S_0 = number of rows
        cj[0] = (float)(2 * PAIRS);
#pragma unroll
        for (int j = 1; j < N; ++j) {
            float2 acc = a2[0][j - 1];
#pragma unroll
            for (int h = 1; h < PAIRS; ++h) acc = fadd2(acc, a2[h][j - 1]);
            cj[j] = acc.x + acc.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            float2 acc = fmul2(a2[0][K - 1], a2[0][K - 1 + j]);
#pragma unroll
            for (int h = 1; h < PAIRS; ++h) acc = ffma2(a2[h][K - 1], a2[h][K - 1 + j], acc);
            cj[j] = acc.x + acc.y;
        }
    }
    float own;
    warp_allsum<N>(cj, lane, own);
    constexpr int LPC = AllSum<N>::LANES_PER_COLUMN;
    if (lane % LPC == 0 && lane / LPC < N) srow[K + lane / LPC] = own;
    const float nrk = -rcp_approx(cj[0]);  // -1 / S_k
#pragma unroll
    for (int j = 1; j < N; ++j) cj[j] *= nrk;  // -(2 * dot / u_length_squared) of bmfr.cl:650
#pragma unroll
    for (int j = 1; j < N; ++j) {
        const float c = cj[j];
        const float2 c2 = make_float2(c, c);
#pragma unroll
        for (int h = 0; h < PAIRS; ++h) {
            if (K == 0) a2[h][j - 1] = fadd2(a2[h][j - 1], c2);
            else a2[h][K - 1 + j] = ffma2(a2[h][K - 1], c2, a2[h][K - 1 + j]);
        }
    }
}
template <int PAIRS, int K>
struct QrLoop2 {
    static __device__ __forceinline__ void run(float2 (&a2)[PAIRS][BMFR_BUFFER_COUNT - 1], float* srows, int lane) {
        qr_step2<PAIRS, K>(a2, srows + K * BMFR_BUFFER_COUNT, lane);
        QrLoop2<PAIRS, K + 1>::run(a2, srows, lane);
    }
};
template <int PAIRS>
struct QrLoop2<PAIRS, BMFR_FEATURES> {
    static __device__ __forceinline__ void run(float2 (&)[PAIRS][BMFR_BUFFER_COUNT - 1], float*, int) {}
};


// --------------------------------------------------------------------------------------------
// fit_qr_kernel: persistent CTAs of four warps, four CTAs per SM.
//
// A CTA walks over blocks: its first block is blockIdx.x, the following ones are drawn from a global
// counter (a block ahead while many are left, at the last moment near the end).
//   warp w, lane l : rows (x_in = l, y_in = 8w .. 8w+7) of the block, eight rows per thread (as four
//       packed fp32 pairs), so one warp-wide reduction serves 256 matrix rows.  Thread 0 draws the next
//       block, publishes its index and starts the TMA loads of its three input tiles while the current
//       block is factored (border blocks, which need mirroring, are loaded pixel by pixel).  Per block:
//       block min/max through one CTA barrier, scaling + noise, level 1 of the TSQR (this warp's 256
//       rows -> one 10x13 triangle), written to global memory (it stays in L2).
//   after its last block the CTA runs level 2 for the blocks it factored: 40 stacked rows -> R, back-
//       substitution (bmfr.cl:659-699), weights.  Nothing crosses CTAs.  Level 2 used to live in a fifth,
//       concurrent "solver" warp fed through a shared-memory ring: its unrolled code competed with level 1
//       for the instruction cache (`no_instruction` was the top stall) and for issue slots, and the 160-
//       thread CTAs fit only three to an SM.
// --------------------------------------------------------------------------------------------
#define QR_COMPUTE_WARPS 4
#define QR_THREADS (QR_COMPUTE_WARPS * 32)
#define QR_ROWS 8
#define QR_TRI (BMFR_FEATURES * BMFR_BUFFER_COUNT)  // floats of one triangle, stored as a full 10x13
#ifndef QR_MINE
#define QR_MINE 128  // blocks a CTA collects before it runs level 2 on them (normally: all its blocks, once, at the end)
#endif
#define QR_TRI_G 136  // floats per level-1 triangle in global memory: four of them are a whole number of 128-byte lines
// One 32x32-pixel tile of an interleaved-RGB image is 96 floats per row.  TMA wants the innermost
// start coordinate on a 16-byte boundary; a tile starts at pixel x0 (even), i.e. at float 3*x0 = 0 or 2
// (mod 4), so the box is 100 floats wide, starts at the aligned-down coordinate and the reader skips
// `shift` = 0 or 2 floats.
#define QR_TILE_W 100
#define QR_TILE_FLOATS (32 * QR_TILE_W)
#define QR_TILE_BYTES (QR_TILE_FLOATS * 4)

struct QrShared {
    float stage[3][32][QR_TILE_W];               // TMA landing zone: normals, positions, accumulated colour of the next block
    float minmax[2][QR_COMPUTE_WARPS][2 * BMFR_FEATURES_SCALED];  // double-buffered by block parity
    float fin[QR_COMPUTE_WARPS][2 * QR_TRI];     // level-2 triangles: two blocks per solving warp
    int mine[QR_MINE];                           // blocks this CTA has factored and not yet solved
    unsigned long long data_full;                // mbarrier: the tiles of the next block have landed / its index is published
    int blk[2];                                  // block index of the next iteration, by iteration parity (dynamic schedule)
};

// The three tensor maps of a frame (2-D tensors [rows][W*3] of floats, box 100 x 32) and whether the
// TMA path can be used at all (W % 4 == 0, 16-byte aligned bases, driver entry point found).
struct QrMaps {
    CUtensorMap normals, positions, colour;
    int use_tma;
};

// The 32 source indices mirror(x0 .. x0 + 31) of a block edge (bmfr.cl:209-216, 314-316) always lie within 32 consecutive
// pixels: a block inside the image is its own source, a block that straddles or lies beyond a border folds back onto
// the 32 pixels next to that border.  Returns the first pixel of that 32-pixel window (inside [0, size - 32]).
__device__ __forceinline__ int qr_box_origin(int x0, int size) {
    int lo;
    if (x0 >= 0 && x0 + 32 <= size) lo = x0;
    else if (x0 < 0) lo = (x0 + 31 >= 0) ? 0 : -x0 - 32;
    else lo = (x0 < size) ? min(x0, 2 * size - x0 - 32) : 2 * size - x0 - 32;
    return min(max(lo, 0), size - 32);
}
// Every block is therefore fetched as three 32x32-pixel TMA tiles at (ox, oy) and read with mirrored indices — as
// long as this context holds the window's rows; otherwise (a strip edge) the block is loaded pixel by pixel, which
// also reports rows the strip does not hold.
__device__ __forceinline__ bool qr_block_box(const KParams& P, int bx, int by, int& ox, int& oy) {
    ox = qr_box_origin(bx * 32 - 16 + P.off_x, P.W);
    oy = qr_box_origin(by * 32 - 16 + P.off_y, P.H);
    return oy >= P.row0 && oy + 32 <= P.row1;
}
// One thread: arm the barrier and start the three tile loads of block (bx, by).  part 0: everything;
// part 1: the caller's inputs only (normals, positions — they do not depend on this frame's
// reprojection and can be requested before the grid dependency is resolved); part 2: the colour tile.
template <class SH>
__device__ __forceinline__ void qr_prefetch(const KParams& P, const QrMaps& M, SH& sh, int ox, int oy, int part = 0) {
    const int c0 = (ox * 3) & ~3, c1 = oy - P.row0;
    if (part != 2) {
        mbar_expect_tx(&sh.data_full, 3 * QR_TILE_BYTES);
#if BMFR_L2_HINTS
        tma_load_tile_hint(&sh.stage[0][0][0], &M.normals, c0, c1, &sh.data_full, BMFR_L2_KEEP);  // the post pass reads them again
        tma_load_tile_hint(&sh.stage[1][0][0], &M.positions, c0, c1, &sh.data_full, BMFR_L2_KEEP);
#else
        tma_load_tile(&sh.stage[0][0][0], &M.normals, c0, c1, &sh.data_full);
        tma_load_tile(&sh.stage[1][0][0], &M.positions, c0, c1, &sh.data_full);
#endif
    }
    if (part != 1) tma_load_tile(&sh.stage[2][0][0], &M.colour, c0, c1, &sh.data_full);
}

__device__ __forceinline__ int qr_block_of_draw(int i, int nblocks, int blocks_x, int frame);
// One thread: draw the block of iteration it + 1 (sh.blk holds the block index, or >= nblocks when the
// frame is exhausted), start its tile loads or arrive plainly.
template <class SH>
__device__ __forceinline__ void qr_draw_next(const KParams& P, const QrMaps& M, SH& sh, int it, int nblocks, int stride) {
    const int draw = stride + atomicAdd(P.block_counter, 1);
    const int nl = draw < nblocks ? qr_block_of_draw(draw, nblocks, P.blocks_x, P.frame) : nblocks;
    sh.blk[(it + 1) & 1] = nl;
    int ox, oy;
    if (nl < nblocks && qr_block_box(P, nl % P.blocks_x, P.by0 + nl / P.blocks_x, ox, oy) && M.use_tma) qr_prefetch(P, M, sh, ox, oy);
    else mbar_arrive(&sh.data_full);
}

// The second half of qr_draw_next() for a draw that was issued earlier (fit_gram_kernel): publish the block, start its loads.
template <class SH>
__device__ __forceinline__ void qr_start_next(const KParams& P, const QrMaps& M, SH& sh, int it, int nblocks, int draw) {
    const int nl = draw < nblocks ? qr_block_of_draw(draw, nblocks, P.blocks_x, P.frame) : nblocks;
    sh.blk[(it + 1) & 1] = nl;
    int ox, oy;
    if (nl < nblocks && qr_block_box(P, nl % P.blocks_x, P.by0 + nl / P.blocks_x, ox, oy) && M.use_tma) qr_prefetch(P, M, sh, ox, oy);
    else mbar_arrive(&sh.data_full);
}
#ifndef BMFR_GRAM_EARLY_DRAW
#define BMFR_GRAM_EARLY_DRAW 1
#endif

#ifndef BMFR_QR_MIN_BLOCKS
#define BMFR_QR_MIN_BLOCKS 4
#endif

// Level 2 of the TSQR and the back-substitution for TWO blocks by one warp, written
// for code size: this runs once per block, so its instructions are fetched cold, and a fully unrolled
// reflector chain (tens of KB) is bound by instruction fetch, not by arithmetic.
//   Half-warp h = lane >> 4 works on block blk (negative: none); its lane j < 13 owns COLUMN j of the 40
//   stacked level-1 rows (row (w,k) = S_kj / sqrt(S_kk) for j >= k, zero left of the diagonal).  A
//   reflector is then: broadcast column k (40 indexed shuffles), one in-lane dot product per column, one
//   in-lane update — no reduction across lanes, and the k loop stays rolled because only the shuffle
//   source lane depends on k.
//   tri = the block's four level-1 triangles in global memory (written by this CTA), fin = 130 floats of
//   shared memory for this half-warp.
__device__ __forceinline__ void qr_solve_columns(const KParams& P, const float* __restrict__ tri, float* __restrict__ fin, int blk,
                                                 int lane) {
    constexpr int NROWS = QR_COMPUTE_WARPS * BMFR_FEATURES;
    const int j = lane & 15, hb = lane & 16;
    const bool owner = blk >= 0 && j < BMFR_BUFFER_COUNT;
    float col[NROWS];
#pragma unroll
    for (int r = 0; r < NROWS; ++r) {
        const int w = r / BMFR_FEATURES, k = r % BMFR_FEATURES;
        const float v = owner ? __ldcg(tri + w * QR_TRI_G + k * BMFR_BUFFER_COUNT + j) : 0.f;
        const float diag = __shfl_sync(0xffffffffu, v, hb + k);  // S_kk of this row
        col[r] = (owner && j >= k) ? v * rsqrt_approx(diag) : 0.f;  // (entries left of the diagonal were never written)
    }
#pragma unroll 1
    for (int k = 0; k < BMFR_FEATURES; ++k) {
        float ck[NROWS];
#pragma unroll
        for (int r = 0; r < NROWS; ++r) ck[r] = __shfl_sync(0xffffffffu, col[r], hb + k);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < NROWS; ++r) acc[r & 3] = fmaf(ck[r], col[r], acc[r & 3]);
        const float S = (acc[0] + acc[1]) + (acc[2] + acc[3]);  // S_j = a_k . a_j
        const float Sk = __shfl_sync(0xffffffffu, S, hb + k);
        if (owner && j >= k) fin[k * BMFR_BUFFER_COUNT + j] = S;  // unnormalised row k of R
        const float c = (j > k) ? -S * rcp_approx(Sk) : 0.f;     // columns <= k are finished
#pragma unroll
        for (int r = 0; r < NROWS; ++r) col[r] = fmaf(ck[r], c, col[r]);
    }
    __syncwarp();
    // back-substitution, bmfr.cl:659-692, one colour channel per lane (j < 3), R read as shared-memory
    // broadcasts.  Row i of R is S_ij / sqrt(S_ii); the square root cancels in R x = rhs.
    if (blk >= 0 && j < 3) {
        float x[BMFR_FEATURES];
#pragma unroll
        for (int i = BMFR_FEATURES - 1; i >= 0; --i) {
            float a = fin[i * BMFR_BUFFER_COUNT + BMFR_FEATURES + j];
#pragma unroll
            for (int jj = i + 1; jj < BMFR_FEATURES; ++jj) a = fmaf(-fin[i * BMFR_BUFFER_COUNT + jj], x[jj], a);
            x[i] = a * rcp_approx(fin[i * BMFR_BUFFER_COUNT + i]);
        }
        float* wout = P.weights + (size_t)(P.by0 * P.blocks_x + blk) * BMFR_FEATURES * 3 + j;  // bmfr.cl:694-699
#pragma unroll
        for (int i = 0; i < BMFR_FEATURES; ++i) wout[i * 3] = x[i];
    }
    __syncwarp();  // fin is reused by this half-warp's next block
}

// Level 2 for the `count` blocks listed in sh.mine, whose triangles this CTA wrote to
// global memory.  Entry e goes to half-warp (e / 4) % 2 of warp e % 4, so a short list still spreads over
// all four warps.  Nothing crosses CTAs: two CTA barriers are all the synchronisation there is.
__device__ __noinline__ void qr_solve_mine(const KParams& P, QrShared& sh, int count, int warp, int lane) {
    __syncthreads();  // the triangles (global stores of this CTA) and the list are complete
    for (int t = 0; t * 8 + warp < count; ++t) {
        const int e = t * 8 + (lane >> 4) * 4 + warp;
        const int blk = e < count ? sh.mine[e] : -1;
        qr_solve_columns(P, P.tri + (size_t)(blk >= 0 ? blk : 0) * QR_COMPUTE_WARPS * QR_TRI_G, &sh.fin[warp][0] + (lane >> 4) * QR_TRI,
                         blk, lane);
    }
    __syncthreads();  // the list may be refilled
}

// Optional phase timers (-DBMFR_QR_TIMING, tuning builds only): clock64 stamps of CTA 0's first
// thread, read back with bmfr_debug_qr_timing(); per CTA start / end of level 1 / end of level 2.
#ifdef BMFR_QR_TIMING
__device__ long long g_qr_timing[512];
#define QR_STAMP(base, it, k)                                                              \
    do {                                                                                   \
        if (blockIdx.x == 0 && lane == 0 && (it) < 8) g_qr_timing[(base) + (it) * 8 + (k)] = clock64(); \
    } while (0)
__device__ long long g_qr_cta[4096];  // per CTA: start, end of level 1, end of level 2 (globaltimer ns), SM id
__device__ __forceinline__ long long qr_globaltimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define QR_CTA_STAMP(k)                                                                         \
    do {                                                                                        \
        if (lane == 0 && blockIdx.x < 1024) g_qr_cta[blockIdx.x * 4 + (k)] = qr_globaltimer(); \
    } while (0)
extern "C" int bmfr_debug_qr_timing(long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, g_qr_timing, sizeof(long long) * (n < 512 ? n : 512));
}
extern "C" int bmfr_debug_qr_cta(long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, g_qr_cta, sizeof(long long) * (n < 4096 ? n : 4096));
}
#else
#define QR_STAMP(base, it, k) do { } while (0)
#define QR_CTA_STAMP(k) do { } while (0)
#endif

// Draw order -> block: the last block row first (with the first row right after it: on a full frame
// both need mirroring and take longer), so that the blocks drawn last are cheap interior ones.
#ifndef BMFR_FIT_REVERSE
#define BMFR_FIT_REVERSE 1
#endif
__device__ __forceinline__ int qr_block_of_draw(int i, int nblocks, int blocks_x, int frame) {
#if BMFR_FIT_REVERSE
    // against the reprojection's sweep: it finished where this starts, so the first wave's tiles — 444 CTAs asking for
    // 17 MB at once — are the likeliest to be in L2 still (measured: 36.6 -> 35.5 us)
    return sweep_down(frame) ? nblocks - 1 - i : i;
#else
    return i < blocks_x ? nblocks - blocks_x + i : i - blocks_x;
#endif
}
#ifndef BMFR_QR_LAZY_DIV
#define BMFR_QR_LAZY_DIV 1  // draws become late once fewer than gridDim.x / LAZY_DIV blocks are left (0: never)
#endif

template <bool STRIP>
__global__ void __launch_bounds__(QR_THREADS, BMFR_QR_MIN_BLOCKS) fit_qr_kernel(const __grid_constant__ KParams P,
                                                                                const __grid_constant__ QrMaps M) {
    // (the 128-byte alignment the TMA destination needs comes from the declaration: rounding the address
    // up through an integer would hide the address space from the compiler, and every access to `sh`
    // would become a generic LD / ST instead of LDS / STS)
    extern __shared__ __align__(128) unsigned char qr_smem[];
    QrShared& sh = *reinterpret_cast<QrShared*>(qr_smem);
    const int tid = threadIdx.x, lane = tid & 31;
    // broadcast from lane 0 so that the compiler knows that branches on `warp` are warp-uniform (otherwise
    // the shuffles behind them get a second, divergence-safe copy and the kernel outgrows the instruction
    // cache)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    constexpr int NSC = BMFR_FEATURES_SCALED, NNS = BMFR_FEATURES_NOT_SCALED, ROWS = QR_ROWS;
    // Block schedule: the first block of a CTA is blockIdx.x, the following ones come from a global
    // counter one block ahead (thread 0 draws the index, publishes it and starts the loads while the
    // CTA factors the current block), so CTAs that drew cheap blocks or fast SMs take more.
    const int nblocks = P.blocks_x * (P.by1 - P.by0);
    const int stride = gridDim.x;
    if ((int)blockIdx.x >= nblocks) return;
    const int first = qr_block_of_draw(blockIdx.x, nblocks, P.blocks_x, P.frame);

    if (tid == 0) {
        mbar_init(&sh.data_full, 1);
        mbar_fence_init();
    }
    __syncthreads();
    // before the grid dependency: the first block's normals / positions tiles (the caller's inputs)
    int fox, foy;
    const bool first_by_tma = qr_block_box(P, first % P.blocks_x, P.by0 + first / P.blocks_x, fox, foy) && M.use_tma;
    if (tid == 0 && first_by_tma) qr_prefetch(P, M, sh, fox, foy, 1);
    pdl_wait();     // the reprojection of this frame is complete (accumulated colour, noise tile, block counter)
    pdl_trigger();  // after the wait, so that completion of everything before this grid is transitive for the post pass
    stamp_begin(P, 1);

    if (warp == 0) {
        QR_CTA_STAMP(0);
#ifdef BMFR_QR_TIMING
        if (lane == 0 && blockIdx.x < 1024) {
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            g_qr_cta[blockIdx.x * 4 + 3] = smid;
        }
#endif
    }
    // ---------------- level 1, block after block ----------------
    // data_full completes once per iteration: thread 0 arrives on it after publishing sh.blk[it & 1],
    // with the TMA byte count when the block is fetched as tiles, plainly otherwise.
    if (tid == 0) {
        sh.blk[0] = first;
        if (first_by_tma) qr_prefetch(P, M, sh, fox, foy, 2);
        else mbar_arrive(&sh.data_full);
    }
    int it = 0;
    int mine = 0;  // blocks of this CTA whose triangles wait in global memory for level 2
    for (;; ++it) {
        mbar_wait_hot(&sh.data_full, it & 1);
        const int local = sh.blk[it & 1];
        if (local >= nblocks) break;
        const int group = P.by0 * P.blocks_x + local;
        const int bx = local % P.blocks_x, by = P.by0 + local / P.blocks_x;

        // a[s][c-1] = column c of row (x_in = lane, y_in = 8 warp + s): the 12 non-constant K1 values
        // (bmfr.cl:448-453), NaN -> 0 (bmfr.cl:468-469)
        float a[ROWS][BMFR_BUFFER_COUNT - 1];
        if (warp == 0) QR_STAMP(0, it, 0);
        int ox, oy;
        if (qr_block_box(P, bx, by, ox, oy) && M.use_tma) {
            // this lane's pixel column and this thread's eight rows inside the 32x32 window, mirrored like bmfr.cl:314-316
            const int col = ((ox * 3) & 3) + 3 * (mirror_index(bx * 32 + lane - 16 + P.off_x, P.W) - ox);
            const int y_in = by * 32 + warp * ROWS - 16 + P.off_y;  // shift + this lane's pixel
            // NaN -> 0 costs two instructions per value; NaNs are rare, so the values are only probed
            // (one predicate-accumulating compare each) and scrubbed in a cold path if any lane saw one
            // (a squared NaN is a NaN and scrubs to 0 = the square of the scrubbed value).
            bool bad = false;
#pragma unroll
            for (int s = 0; s < ROWS; ++s) {
                float v[9];
#pragma unroll
                for (int c = 0; c < 9; ++c) {
                    v[c] = sh.stage[c / 3][mirror_index(y_in + s, P.H) - oy][col + c % 3];
                    bad = bad || (v[c] != v[c]);
                }
                a[s][0] = v[0]; a[s][1] = v[1]; a[s][2] = v[2];
                a[s][3] = v[3]; a[s][4] = v[4]; a[s][5] = v[5];
                a[s][6] = v[3] * v[3]; a[s][7] = v[4] * v[4]; a[s][8] = v[5] * v[5];
                a[s][9] = v[6]; a[s][10] = v[7]; a[s][11] = v[8];
            }
            if (__any_sync(0xffffffffu, bad)) {
#pragma unroll
                for (int s = 0; s < ROWS; ++s)
#pragma unroll
                    for (int c = 0; c < BMFR_BUFFER_COUNT - 1; ++c) a[s][c] = scrub_nan(a[s][c]);
            }
        } else {  // the window leaves the rows this strip holds (or no tensor maps): pixel by pixel
            const int x = mirror_index(bx * 32 + lane - 16 + P.off_x, P.W);
#pragma unroll
            for (int s = 0; s < ROWS; ++s) {
                const int y = mirror_index(by * 32 + warp * ROWS + s - 16 + P.off_y, P.H);
                if (STRIP && (y < P.row0 || y >= P.row1)) {
                    *P.oob_flag = 1;
#pragma unroll
                    for (int c = 0; c < BMFR_BUFFER_COUNT - 1; ++c) a[s][c] = 0.f;
                    continue;
                }
                const unsigned int lp = pix_index(P, x, y);
                const f3 n = load_f3(P.cur_normals, lp);
                const f3 p = load_f3(P.cur_positions, lp);
                const f3 col = load_f3(P.cur_noisy_acc, lp);
                const float px = scrub_nan(p.x), py = scrub_nan(p.y), pz = scrub_nan(p.z);
                a[s][0] = scrub_nan(n.x); a[s][1] = scrub_nan(n.y); a[s][2] = scrub_nan(n.z);
                a[s][3] = px; a[s][4] = py; a[s][5] = pz;
                a[s][6] = px * px; a[s][7] = py * py; a[s][8] = pz * pz;
                a[s][9] = scrub_nan(col.x); a[s][10] = scrub_nan(col.y); a[s][11] = scrub_nan(col.z);
            }
        }
        if (warp == 0) QR_STAMP(0, it, 1);

        // (i) block min / max of the six scaled features, bmfr.cl:511-535 (exact, so order-free)
#pragma unroll
        for (int f = 0; f < NSC; ++f) {
            const int c = NNS - 1 + f;
            float lo = a[0][c], hi = a[0][c];
#pragma unroll
            for (int s = 1; s < ROWS; ++s) {
                lo = fminf(lo, a[s][c]);
                hi = fmaxf(hi, a[s][c]);
            }
            const float wlo = warp_min(lo), whi = warp_max(hi);
            if (lane == 0) {
                sh.minmax[it & 1][warp][2 * f] = wlo;
                sh.minmax[it & 1][warp][2 * f + 1] = whi;
            }
        }
        if (warp == 0) QR_STAMP(0, it, 2);
        __syncthreads();  // the per-warp extrema are visible, and every thread is done with the stage
        // Thread 0 draws the next block, publishes it and starts its loads: right here, a whole block
        // ahead, while many blocks are left; but once less than a round is left an early draw would park a
        // block behind this CTA's current one while other CTAs run dry, so the draw moves to the end of
        // the iteration (the last blocks go to whoever is free first, at the price of an exposed load).
        bool late_draw = false;
        if (tid == 0) {
            sh.mine[mine] = local;
            if (BMFR_QR_LAZY_DIV > 0)
                late_draw = nblocks - stride - *(volatile int*)P.block_counter < stride / (BMFR_QR_LAZY_DIV > 0 ? BMFR_QR_LAZY_DIV : 1);
            if (!late_draw) qr_draw_next(P, M, sh, it, nblocks, stride);
        }
        // every warp finishes the reduction itself (lane f < 6 owns feature f) and shares the result
        // by shuffle: one block-wide barrier per block instead of two
        float mn[NSC], inv[NSC];
        {
            const int f = lane < NSC ? lane : 0;
            float lo = sh.minmax[it & 1][0][2 * f], hi = sh.minmax[it & 1][0][2 * f + 1];
#pragma unroll
            for (int w = 1; w < QR_COMPUTE_WARPS; ++w) {
                lo = fminf(lo, sh.minmax[it & 1][w][2 * f]);
                hi = fmaxf(hi, sh.minmax[it & 1][w][2 * f + 1]);
            }
            const float iv = scale_factor(lo, hi);
            if (warp == 0 && lane < NSC) {
                P.mins_maxs[(size_t)group * 2 * NSC + 2 * lane] = lo;
                P.mins_maxs[(size_t)group * 2 * NSC + 2 * lane + 1] = hi;
                P.mins_inv[(size_t)group * 2 * NSC + 2 * lane] = lo;
                P.mins_inv[(size_t)group * 2 * NSC + 2 * lane + 1] = iv;
            }
#pragma unroll
            for (int k = 0; k < NSC; ++k) {
                mn[k] = __shfl_sync(0xffffffffu, lo, k);
                inv[k] = __shfl_sync(0xffffffffu, iv, k);
            }
        }
        if (warp == 0) QR_STAMP(0, it, 3);

        // scale (bmfr.cl:538-541), then the first-touch noise on columns 1..9 (bmfr.cl:623-627).  The
        // reference adds a double (NOISE_AMOUNT is a double literal); the tile holds that double
        // rounded to fp32, which changes a sum by at most one ulp in rare ties — below the fit's own
        // rounding.
        float2 a2[ROWS / 2][BMFR_BUFFER_COUNT - 1];  // rows (2h, 2h+1) packed: FADD2 / FMUL2 / FFMA2 from here on
#pragma unroll
        for (int h = 0; h < ROWS / 2; ++h)
#pragma unroll
            for (int c = 0; c < BMFR_BUFFER_COUNT - 1; ++c) a2[h][c] = make_float2(a[2 * h][c], a[2 * h + 1][c]);
#pragma unroll
        for (int f = 0; f < NSC; ++f) {
            const float2 mn2 = dup2(mn[f]), inv2 = dup2(inv[f]);
#pragma unroll
            for (int h = 0; h < ROWS / 2; ++h) a2[h][NNS - 1 + f] = fmul2(fsub2(a2[h][NNS - 1 + f], mn2), inv2);
        }
        {
            const float4* nz4 = reinterpret_cast<const float4*>(P.noise_f) + (size_t)warp * (BMFR_FEATURES - 1) * 2 * 32 + lane;
#pragma unroll
            for (int c = 0; c < BMFR_FEATURES - 1; ++c) {
#pragma unroll
                for (int q = 0; q < ROWS / 4; ++q) {
                    const float4 nz = __ldg(nz4 + (c * 2 + q) * 32);
                    a2[2 * q][c] = fadd2(a2[2 * q][c], make_float2(nz.x, nz.y));
                    a2[2 * q + 1][c] = fadd2(a2[2 * q + 1][c], make_float2(nz.z, nz.w));
                }
            }
        }

        // (ii) level 1 of the TSQR: this warp's 256 rows -> one 10x13 triangle (global memory, stays in L2)
        if (warp == 0) QR_STAMP(0, it, 4);
        if (warp == 0) QR_STAMP(0, it, 5);
        QrLoop2<ROWS / 2, 0>::run(a2, P.tri + ((size_t)local * QR_COMPUTE_WARPS + warp) * QR_TRI_G, lane);
        if (++mine == QR_MINE) {  // the list is full (more than 128 blocks per CTA: no frame up to 8K on a B200 gets here)
            qr_solve_mine(P, sh, mine, warp, lane);
            mine = 0;
        }
        if (tid == 0 && late_draw) qr_draw_next(P, M, sh, it, nblocks, stride);
        if (warp == 0) QR_STAMP(0, it, 6);
    }
    if (warp == 0) QR_CTA_STAMP(1);
    // ---------------- level 2 + back-substitution of this CTA's blocks ----------------
    qr_solve_mine(P, sh, mine, warp, lane);
    if (warp == 0) QR_CTA_STAMP(2);
    stamp_end(P, 1);
}


// ================================================================================================
// fit_gram_kernel: the fit through the block's Gram matrix (the default; fit_qr_kernel above stays selectable with
// bmfr_params.fit_method = BMFR_FIT_TSQR).
//
// The two-level QR above spends a third of its issue slots on ten DEPENDENT warp all-reductions per block (one per
// reflector) and another third on the eliminations; profiles/r01_v7_ncu_full.md has it issue-bound.  The least-squares
// solution only needs R^T R = A^T A and A^T y, i.e. the 13x13 Gram matrix G of [1 | features | colour]: 90 independent
// dot products per thread (FFMA2 on packed row pairs, no dependence between them), ONE reduction over the warp for all
// of them (a shared-memory transpose: 90 stores, 24 128-bit loads and a pairwise add tree per lane), the four warps'
// totals and the Cholesky factorisation + substitutions in fp64 by half a warp per block after the CTA's last block.
//
// Accuracy: the normal equations square the condition number (about 500 here: planar geometry makes positions and
// their squares collinear up to the 1e-2 regularisation noise, bmfr.cl:623-627), so the columns are first centred on
// the block means (the constant column stays in G, so centring by a rounded mean is still exact algebra; the
// intercept is recovered after the solve).  Measured on synth-v1 blocks against an fp64 least-squares fit, worst
// relative error of the fitted colour: reference-order fp32 Householder 1.3e-5, this scheme 1.9e-5, uncentred fp32
// normal equations 3e-4 (scripts/gram_accuracy.py); parity tests hold the frame to 1e-3 / 60 dB as before.
// ================================================================================================
// Scratch per (block, warp): QR_TRI_G floats — the Gram entries (90 for the default list), then the block means of the
// non-constant columns (written by warp 0).
#define GR_RED_W 36     // floats per row of the transpose buffer: 16-byte aligned rows, conflict-free 128-bit reads
// CTAs per SM.  Measured at 1080p (profiles/r02_f_*): 3 -> 35.2 us, 4 (with 16-entry reduction rounds, 128 registers) ->
// 45.3 us: a CTA's first block is its most expensive one (cold landing zone, cold instruction cache) and a fourth CTA
// takes the L1 away, so fewer, longer-lived CTAs win.
#ifndef BMFR_GRAM_MIN_BLOCKS
#define BMFR_GRAM_MIN_BLOCKS 3
#endif
// Gram entries reduced per round trip through the transpose buffer.  32 uses every lane for the row sums; 16 halves the
// buffer (2.3 KB per warp), which is what lets a fourth CTA fit on an SM next to the 38 KB landing zones.
#ifndef GR_CHUNK
#define GR_CHUNK (BMFR_GRAM_MIN_BLOCKS >= 4 ? 16 : 32)
#endif

struct GramShared {
    float stage[3][32][QR_TILE_W];                  // TMA landing zone (as QrShared)
    float red[QR_COMPUTE_WARPS][GR_CHUNK][GR_RED_W];  // per-warp transpose buffer: [entry of the chunk][lane]; between blocks the solver's fp64 workspace
    float part[2][QR_COMPUTE_WARPS][24];            // per warp: min (6), max (6), column sums (12); double-buffered by block parity
    int mine[QR_MINE];
    unsigned long long data_full;
    int blk[2];
};
static_assert(2 * BMFR_BUFFER_COUNT * BMFR_FEATURES * sizeof(double) <= GR_CHUNK * GR_RED_W * sizeof(float), "solver workspace (largest list) fits the transpose buffer");
static_assert(GR_CHUNK >= BMFR_BUFFER_COUNT - 1 && GR_CHUNK <= 32, "the column sums of a block go through the buffer in one round");

// Sum of the 32 floats of one row of the transpose buffer (eight 128-bit loads, pairwise tree).
#ifndef BMFR_GRAM_PACKED_SUM
#define BMFR_GRAM_PACKED_SUM 1  // the tree on packed pairs: 15 FADD2 + 1 FADD instead of 31 FADD
#endif
#ifndef BMFR_GRAM_INTERIOR_LOADS
#define BMFR_GRAM_INTERIOR_LOADS 1
#endif
__device__ __forceinline__ float gram_row_sum(const float* __restrict__ row) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
#if BMFR_GRAM_PACKED_SUM
    float2 t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 v = r4[i];
        t[i] = fadd2(make_float2(v.x, v.y), make_float2(v.z, v.w));
    }
    const float2 u = fadd2(fadd2(fadd2(t[0], t[1]), fadd2(t[2], t[3])), fadd2(fadd2(t[4], t[5]), fadd2(t[6], t[7])));
    return u.x + u.y;
#else
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 v = r4[i];
        t[i] = (v.x + v.y) + (v.z + v.w);
    }
    return ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
#endif
}

__device__ __forceinline__ double shfl_f64(double v, int src) {
    const int lo = __shfl_sync(0xffffffffu, __double2loint(v), src), hi = __shfl_sync(0xffffffffu, __double2hiint(v), src);
    return __hiloint2double(hi, lo);
}
// 1 / d in fp64 from the fp32 approximation (relative error 2^-22) and one Newton step (-> 2^-43; the Gram entries carry
// 2^-24).  No DDIV sequence: this code runs once per block and is fetched cold, so it is written for size and for a short
// dependent chain.
__device__ __forceinline__ double rcp_f64(double d) {
    const double r = (double)rcp_approx((float)d);
    return fma(fma(-d, r, 1.0), r, r);
}
// index of G_ij (i <= j, (i, j) != (0, 0)) in the scratch order
template <int NCOL>
__device__ __forceinline__ int gram_index(int i, int j) {
    return i == 0 ? j - 1 : NCOL + (i - 1) * NCOL - ((i - 1) * (i - 2)) / 2 + (j - i);
}

// Factorisation + substitution for the block `blk` by half a warp (hb = lane & 16): lane i < NC owns row i of G.
// G = L D L^T with a unit lower triangle (no square roots, and the pivot row's entries are broadcast while its reciprocal
// is still being refined); carried through the three colour rows it leaves them as D^-1 L^-1 [A^T y], so the weights are
// the solution of the unit triangular system L^T x = that — no division (bmfr.cl:659-692 does the same job on R).
template <int FS>
__device__ __forceinline__ void gram_solve(const KParams& P, double* __restrict__ fin, int blk, int lane) {
    constexpr int NF = FeatureSet<FS>::F, NC = NF + 3, NCOL = NC - 1, ENTRIES = NCOL + NCOL * (NCOL + 1) / 2;
    const int i = lane & 15, hb = lane & 16;
    const bool owner = blk >= 0 && i < NC;
    const float* sc = P.tri + (size_t)(blk >= 0 ? blk : 0) * QR_COMPUTE_WARPS * QR_TRI_G;
    double g[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        double v = 0.0;
        if (owner && (i | j) != 0) {
            const int e = gram_index<NCOL>(i < j ? i : j, i < j ? j : i);
#pragma unroll
            for (int w = 0; w < QR_COMPUTE_WARPS; ++w) v += (double)__ldcg(sc + w * QR_TRI_G + e);  // the four warps, in fp64, fixed order
        }
        g[j] = v;
    }
    if (i == 0) g[0] = (double)BMFR_BLOCK_PIXELS;
    if (!owner) {  // idle lanes: an identity row keeps every operation below finite
#pragma unroll
        for (int j = 0; j < NC; ++j) g[j] = (j == i) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int k = 0; k < NF; ++k) {  // right-looking elimination on the first NF pivots
        double pr[NC];              // the pivot row (= pivot column, G is symmetric) as lane k holds it now
#pragma unroll
        for (int j = k; j < NC; ++j) pr[j] = shfl_f64(g[j], hb + k);
        const double lik = g[k] * rcp_f64(pr[k]);  // L_ik for i > k (lane k itself: 1; lanes above: stale, never read)
#pragma unroll
        for (int j = k + 1; j < NC; ++j) g[j] = fma(-lik, pr[j], g[j]);
        g[k] = lik;
    }
    // L (rows 0..NF-1) and the colour rows -> shared memory, then one colour channel per lane solves L^T x = row (NF + channel)
    if (owner) {
#pragma unroll
        for (int k = 0; k < NF; ++k) fin[i * NF + k] = g[k];
    }
    __syncwarp();
    if (blk >= 0 && i < 3) {
        double x[NF];
#pragma unroll
        for (int r = 0; r < NF; ++r) x[r] = fin[(NF + i) * NF + r];
#pragma unroll
        for (int r = NF - 1; r > 0; --r) {  // column-oriented: x_r is final, every row above it takes its share at once
#pragma unroll
            for (int q = 0; q < r; ++q) x[q] = fma(-fin[r * NF + q], x[r], x[q]);
        }
        // the columns were centred: y - m_y = x_0 + sum_j x_j (a_j - m_j)  ->  intercept of the uncentred model
        const float* mean = sc + ENTRIES;  // block means of the non-constant columns, written by warp 0
        double w0 = x[0] + (double)__ldcg(mean + (NF - 1) + i);
#pragma unroll
        for (int jj = 1; jj < NF; ++jj) w0 -= x[jj] * (double)__ldcg(mean + jj - 1);
        float* wout = P.weights + (size_t)(P.by0 * P.blocks_x + blk) * NF * 3 + i;  // bmfr.cl:694-699
        wout[0] = (float)w0;
#pragma unroll
        for (int r = 1; r < NF; ++r) wout[r * 3] = (float)x[r];
    }
    __syncwarp();  // fin is reused by this half-warp's next block
}

template <int FS, class SH>
__device__ __noinline__ void gram_solve_mine(const KParams& P, SH& sh, int count, int warp, int lane) {
    __syncthreads();  // this CTA's scratch stores and the list are complete
    // fp64 workspace of the two half-warps: this warp's own transpose buffer (idle between blocks)
    double* fin = reinterpret_cast<double*>(&sh.red[warp][0][0]) + (size_t)(lane >> 4) * (BMFR_BUFFER_COUNT * BMFR_FEATURES);
    for (int t = 0; t * 8 + warp < count; ++t) {
        const int e = t * 8 + (lane >> 4) * 4 + warp;
        gram_solve<FS>(P, fin, e < count ? sh.mine[e] : -1, lane);
    }
    __syncthreads();
}

// Tuning builds (-DBMFR_QR_TIMING): globaltimer stamps of every CTA's first thread — [0] start, [1] first tiles landed,
// [2 + 2 it] block it's data ready, [3 + 2 it] block it done (it < 6), [14] level 1 done, [15] solves done.
#ifdef BMFR_QR_TIMING
__device__ long long g_gram_cta[1024 * 16];
#define GRAM_STAMP(k)                                                                                  \
    do {                                                                                               \
        if (tid == 0 && blockIdx.x < 1024 && (k) < 16) g_gram_cta[blockIdx.x * 16 + (k)] = qr_globaltimer(); \
    } while (0)
extern "C" int bmfr_debug_gram_cta(long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, g_gram_cta, sizeof(long long) * (n < 1024 * 16 ? n : 1024 * 16));
}
#else
#define GRAM_STAMP(k) do { } while (0)
#endif

template <bool STRIP, int FS>
__global__ void __launch_bounds__(QR_THREADS, BMFR_GRAM_MIN_BLOCKS) fit_gram_kernel(const __grid_constant__ KParams P,
                                                                                    const __grid_constant__ QrMaps M) {
    extern __shared__ __align__(128) unsigned char qr_smem[];
    GramShared& sh = *reinterpret_cast<GramShared*>(qr_smem);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform (see fit_qr_kernel)
    using FL = FeatureSet<FS>;
    constexpr int NF = FL::F, NSC = FL::NSC, NNS = NF - NSC, ROWS = QR_ROWS, NCOL = NF - 1 + 3;
    constexpr int ENTRIES = NCOL + NCOL * (NCOL + 1) / 2;  // upper triangle of the Gram matrix without G_00 (= 1024): row 0 first, then rows 1..NCOL
    static_assert(ENTRIES + NCOL <= QR_TRI_G && 2 * NSC + NCOL <= 24, "scratch row / per-warp partials");
    const int nblocks = P.blocks_x * (P.by1 - P.by0);
    const int stride = gridDim.x;
    if ((int)blockIdx.x >= nblocks) return;
    const int first = qr_block_of_draw(blockIdx.x, nblocks, P.blocks_x, P.frame);
    GRAM_STAMP(0);

    if (tid == 0) {
        mbar_init(&sh.data_full, 1);
        mbar_fence_init();
    }
    __syncthreads();
    int fox, foy;
    const bool first_by_tma = qr_block_box(P, first % P.blocks_x, P.by0 + first / P.blocks_x, fox, foy) && M.use_tma;
    if (tid == 0 && first_by_tma) qr_prefetch(P, M, sh, fox, foy, 1);
    pdl_wait();     // the reprojection of this frame is complete (accumulated colour, noise tile, block counter)
    pdl_trigger();  // after the wait, so that completion of everything before this grid is transitive for the post pass
    stamp_begin(P, 1);
    if (tid == 0) {
        sh.blk[0] = first;
        if (first_by_tma) qr_prefetch(P, M, sh, fox, foy, 2);
        else mbar_arrive(&sh.data_full);
    }
    float* const red = &sh.red[warp][0][0];
    int mine = 0;
    for (int it = 0;; ++it) {
        mbar_wait_hot(&sh.data_full, it & 1);
        const int local = sh.blk[it & 1];
        if (local >= nblocks) break;
        if (it < 6) GRAM_STAMP(2 + 2 * it);
        const int group = P.by0 * P.blocks_x + local;
        const int bx = local % P.blocks_x, by = P.by0 + local / P.blocks_x;
#if BMFR_GRAM_EARLY_DRAW
        // The draw of the next block, issued a phase ahead of its use: the atomic's round trip to the L2 (and, with the lazy
        // draws, the read of the counter in front of it) used to sit on thread 0's path behind the block barrier — the source-
        // level profile had 9.6 % of ALL warp samples there, i.e. 38 % of warp 0's time, with the next block's tiles requested
        // that much later and the other three warps waiting for them at the top of the next iteration.
        int next_draw = 0;
        if (tid == 0) next_draw = stride + atomicAdd(P.block_counter, 1);
#endif

        // a[s][c-1] = column c of row (x_in = lane, y_in = 8 warp + s): the 12 non-constant K1 values (bmfr.cl:448-453), NaN -> 0
        float a[ROWS][NCOL];
        int ox, oy;
        const bool by_tma = qr_block_box(P, bx, by, ox, oy) && M.use_tma;
        if (BMFR_GRAM_INTERIOR_LOADS && by_tma && ox == bx * 32 - 16 + P.off_x && oy == by * 32 - 16 + P.off_y) {
            // the block lies inside the image (all but the border blocks): the window is the block, nothing is mirrored —
            // one base address, 72 loads at immediate offsets (the mirrored path below spends 3 instructions per load
            // on per-row index arithmetic and its branches, ahead of the loads)
            const float* base = &sh.stage[0][warp * ROWS][((ox * 3) & 3) + 3 * lane];
#pragma unroll
            for (int s = 0; s < ROWS; ++s) {
                float v[9];
#pragma unroll
                for (int c = 0; c < 9; ++c) {
                    if (c < 3 && !FL::NORMALS) continue;  // this list does not read the normals
                    v[c] = base[(c / 3) * QR_TILE_FLOATS + s * QR_TILE_W + c % 3];
                }
                feature_columns<FS>(a[s], v, v + 3, v + 6);  // NaNs are found and scrubbed below
            }
        } else if (by_tma) {
            // this lane's pixel column and this thread's eight rows inside the 32x32 window, mirrored like bmfr.cl:314-316
            const int col = ((ox * 3) & 3) + 3 * (mirror_index(bx * 32 + lane - 16 + P.off_x, P.W) - ox);
            const int y_in = by * 32 + warp * ROWS - 16 + P.off_y;
#pragma unroll
            for (int s = 0; s < ROWS; ++s) {
                float v[9];
#pragma unroll
                for (int c = 0; c < 9; ++c) {
                    if (c < 3 && !FL::NORMALS) continue;  // this list does not read the normals
                    v[c] = sh.stage[c / 3][mirror_index(y_in + s, P.H) - oy][col + c % 3];
                }
                feature_columns<FS>(a[s], v, v + 3, v + 6);  // NaNs are found and scrubbed below
            }
        } else {  // the window leaves the rows this strip holds (or no tensor maps): pixel by pixel
            const int x = mirror_index(bx * 32 + lane - 16 + P.off_x, P.W);
#pragma unroll
            for (int s = 0; s < ROWS; ++s) {
                const int y = mirror_index(by * 32 + warp * ROWS + s - 16 + P.off_y, P.H);
                if (STRIP && (y < P.row0 || y >= P.row1)) {
                    *P.oob_flag = 1;
#pragma unroll
                    for (int c = 0; c < NCOL; ++c) a[s][c] = 0.f;
                    continue;
                }
                const unsigned int lp = pix_index(P, x, y);
                const f3 n = FL::NORMALS ? load_f3(P.cur_normals, lp) : make_f3(0.f, 0.f, 0.f);
                const f3 p = load_f3(P.cur_positions, lp);
                const f3 col = load_f3(P.cur_noisy_acc, lp);
                const float nv[3] = {scrub_nan(n.x), scrub_nan(n.y), scrub_nan(n.z)};
                const float pv[3] = {scrub_nan(p.x), scrub_nan(p.y), scrub_nan(p.z)};
                const float cv[3] = {scrub_nan(col.x), scrub_nan(col.y), scrub_nan(col.z)};
                feature_columns<FS>(a[s], nv, pv, cv);
            }
        }

        // a2[h][c]: rows (2h, 2h + 1) of column c + 1 as one register pair (FADD2 / FFMA2 operands from here on)
        float2 a2[ROWS / 2][NCOL];
#pragma unroll
        for (int h = 0; h < ROWS / 2; ++h)
#pragma unroll
            for (int c = 0; c < NCOL; ++c) a2[h][c] = make_float2(a[2 * h][c], a[2 * h + 1][c]);
        if (FL::SQUARES) {  // the squared positions as packed products (the same correctly rounded products, half the instructions)
            constexpr int LIN = FL::NORMALS ? 3 : 0;
#pragma unroll
            for (int h = 0; h < ROWS / 2; ++h)
#pragma unroll
                for (int j = 0; j < 3; ++j) a2[h][LIN + 3 + j] = fmul2(a2[h][LIN + j], a2[h][LIN + j]);
        }
        // (i) for the centring, the sums of all twelve columns: per thread, then over the warp through the transpose buffer.
        // A NaN input (bmfr.cl:448-453 turns it into 0) poisons its column's sum, so the sums double as the NaN test of
        // the thread's 72 tile values; the scrub itself is the rare path.
        float colsum[NCOL];
        bool bad = false;
#pragma unroll
        for (int c = 0; c < NCOL; ++c) {
            const float2 t = fadd2(fadd2(a2[0][c], a2[1][c]), fadd2(a2[2][c], a2[3][c]));
            colsum[c] = t.x + t.y;
            bad = bad || (colsum[c] != colsum[c]);
        }
        if (by_tma && __any_sync(0xffffffffu, bad)) {
#pragma unroll
            for (int c = 0; c < NCOL; ++c) {
#pragma unroll
                for (int h = 0; h < ROWS / 2; ++h) a2[h][c] = make_float2(scrub_nan(a2[h][c].x), scrub_nan(a2[h][c].y));
                const float2 t = fadd2(fadd2(a2[0][c], a2[1][c]), fadd2(a2[2][c], a2[3][c]));
                colsum[c] = t.x + t.y;
            }
        }
        // block min / max of the scaled features (bmfr.cl:511-535; exact, order-free): three-input min / max per thread,
        // redux over the warp
        float* part = &sh.part[it & 1][warp][0];
#pragma unroll
        for (int f = 0; f < NSC; ++f) {
            const int c = NNS - 1 + f;
            float lo = fmin3(a2[0][c].x, a2[0][c].y, a2[1][c].x), hi = fmax3(a2[0][c].x, a2[0][c].y, a2[1][c].x);
            lo = fmin3(lo, a2[1][c].y, a2[2][c].x); hi = fmax3(hi, a2[1][c].y, a2[2][c].x);
            lo = fmin3(lo, a2[2][c].y, a2[3][c].x); hi = fmax3(hi, a2[2][c].y, a2[3][c].x);
            lo = fminf(lo, a2[3][c].y); hi = fmaxf(hi, a2[3][c].y);
            const float wlo = warp_min(lo), whi = warp_max(hi);
            if (lane == 0) {
                part[f] = wlo;
                part[NSC + f] = whi;
            }
        }
#pragma unroll
        for (int c = 0; c < NCOL; ++c) red[c * GR_RED_W + lane] = colsum[c];
        __syncwarp();
        if (lane < NCOL) part[2 * NSC + lane] = gram_row_sum(red + lane * GR_RED_W);
        __syncthreads();  // the per-warp extrema and sums are visible, and every thread is done with the stage
        bool late_draw = false;
        if (tid == 0) {
            sh.mine[mine] = local;
#if BMFR_GRAM_EARLY_DRAW
            qr_start_next(P, M, sh, it, nblocks, next_draw);
#else
            if (BMFR_QR_LAZY_DIV > 0)
                late_draw = nblocks - stride - *(volatile int*)P.block_counter < stride / (BMFR_QR_LAZY_DIV > 0 ? BMFR_QR_LAZY_DIV : 1);
            if (!late_draw) qr_draw_next(P, M, sh, it, nblocks, stride);
#endif
        }
        // every warp finishes the reductions itself: lane f < NSC owns scaled feature f, lane c < NCOL the mean of column c + 1
        float inv[NSC], mean[NCOL];
        {
            const int f = lane < NSC ? lane : 0, c = lane < NCOL ? lane : 0;
            float lo = sh.part[it & 1][0][f], hi = sh.part[it & 1][0][NSC + f], sum = sh.part[it & 1][0][2 * NSC + c];
#pragma unroll
            for (int w = 1; w < QR_COMPUTE_WARPS; ++w) {
                lo = fminf(lo, sh.part[it & 1][w][f]);
                hi = fmaxf(hi, sh.part[it & 1][w][NSC + f]);
                sum += sh.part[it & 1][w][2 * NSC + c];
            }
            const float iv = scale_factor(lo, hi);
            if (warp == 0 && lane < NSC) {
                P.mins_maxs[(size_t)group * 2 * NSC + 2 * lane] = lo;
                P.mins_maxs[(size_t)group * 2 * NSC + 2 * lane + 1] = hi;
                P.mins_inv[(size_t)group * 2 * NSC + 2 * lane] = lo;
                P.mins_inv[(size_t)group * 2 * NSC + 2 * lane + 1] = iv;
            }
            const float m_raw = sum * (1.0f / BMFR_BLOCK_PIXELS);
#pragma unroll
            for (int k = 0; k < NSC; ++k) inv[k] = __shfl_sync(0xffffffffu, iv, k);
#pragma unroll
            for (int k = 0; k < NCOL; ++k) mean[k] = __shfl_sync(0xffffffffu, m_raw, k);
            // the solver needs the means of the columns as the reference defines them (scaled: (a - min) * inv; the noise
            // averages to ~0) to recover the intercept: one per lane, in the scratch of warp 0
            const int fc = min(max(c - (NNS - 1), 0), NSC - 1);
            const float lo_c = __shfl_sync(0xffffffffu, lo, fc), iv_c = __shfl_sync(0xffffffffu, iv, fc);
            if (warp == 0 && lane < NCOL) {
                const bool scaled = lane >= NNS - 1 && lane < NNS - 1 + NSC;
                P.tri[((size_t)local * QR_COMPUTE_WARPS) * QR_TRI_G + ENTRIES + lane] = scaled ? (m_raw - lo_c) * iv_c : m_raw;
            }
        }

        // centre on the raw block means, then scale (bmfr.cl:538-541; (a - min) inv - mean_scaled = (a - mean_raw) inv) and
        // add the first-touch noise on columns 1..F-1 (bmfr.cl:623-627; from the tile's fp32 rounding as in fit_qr_kernel)
#pragma unroll
        for (int c = 0; c < NCOL; ++c) {
            const float2 m2 = dup2(mean[c]);
#pragma unroll
            for (int h = 0; h < ROWS / 2; ++h) a2[h][c] = fsub2(a2[h][c], m2);
        }
        {
            const float4* nz4 = reinterpret_cast<const float4*>(P.noise_f) + (size_t)warp * (NF - 1) * 2 * 32 + lane;
#pragma unroll
            for (int c = 0; c < NF - 1; ++c) {
                const bool scaled = c >= NNS - 1;
                const float2 inv2 = dup2(scaled ? inv[scaled ? c - (NNS - 1) : 0] : 1.f);
#pragma unroll
                for (int q = 0; q < ROWS / 4; ++q) {
                    const float4 nz = __ldg(nz4 + (c * 2 + q) * 32);
                    if (scaled) {
                        a2[2 * q][c] = ffma2(a2[2 * q][c], inv2, make_float2(nz.x, nz.y));
                        a2[2 * q + 1][c] = ffma2(a2[2 * q + 1][c], inv2, make_float2(nz.z, nz.w));
                    } else {
                        a2[2 * q][c] = fadd2(a2[2 * q][c], make_float2(nz.x, nz.y));
                        a2[2 * q + 1][c] = fadd2(a2[2 * q + 1][c], make_float2(nz.z, nz.w));
                    }
                }
            }
        }

        // (ii) this thread's share of the 90 Gram entries, reduced over the warp 32 entries at a time through the transpose
        // buffer; the warp's totals go to the scratch (global memory, stays in L2) for the solver
        float* const out = P.tri + ((size_t)local * QR_COMPUTE_WARPS + warp) * QR_TRI_G;
        auto flush = [&](int chunk, int count) {
            __syncwarp();
            if (lane < count) out[chunk * GR_CHUNK + lane] = gram_row_sum(red + lane * GR_RED_W);
            __syncwarp();
        };
        int e = 0;  // compile-time after unrolling
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {  // row 0: column sums
            float2 acc = a2[0][j];
#pragma unroll
            for (int h = 1; h < ROWS / 2; ++h) acc = fadd2(acc, a2[h][j]);
            red[(e % GR_CHUNK) * GR_RED_W + lane] = acc.x + acc.y;
            if (++e % GR_CHUNK == 0) flush(e / GR_CHUNK - 1, GR_CHUNK);
        }
#pragma unroll
        for (int i = 0; i < NCOL; ++i) {
#pragma unroll
            for (int j = i; j < NCOL; ++j) {
                float2 acc = fmul2(a2[0][i], a2[0][j]);
#pragma unroll
                for (int h = 1; h < ROWS / 2; ++h) acc = ffma2(a2[h][i], a2[h][j], acc);
                red[(e % GR_CHUNK) * GR_RED_W + lane] = acc.x + acc.y;
                if (++e % GR_CHUNK == 0) flush(e / GR_CHUNK - 1, GR_CHUNK);
            }
        }
        if (ENTRIES % GR_CHUNK != 0) flush(ENTRIES / GR_CHUNK, ENTRIES % GR_CHUNK);

        if (++mine == QR_MINE) {
            gram_solve_mine<FS>(P, sh, mine, warp, lane);
            mine = 0;
        }
        if (tid == 0 && late_draw) qr_draw_next(P, M, sh, it, nblocks, stride);
        if (it < 6) GRAM_STAMP(3 + 2 * it);
    }
    GRAM_STAMP(14);
    gram_solve_mine<FS>(P, sh, mine, warp, lane);
    GRAM_STAMP(15);
    stamp_end(P, 1);
}

// --------------------------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------------------------
static bool is_strip(const KParams& P) { return P.row0 != 0 || P.row1 != P.H; }

// Measured at 1080p (profiles/r02_h_*): 51.8 us with the tiles staged by TMA, 51.6 us with per-thread loads — the
// reprojection is bound by its 430 instructions and 40 tap loads per pixel, not by the first round trip.  Kept as a
// build option, off.
#ifndef BMFR_REPROJECT_TMA
#define BMFR_REPROJECT_TMA 0
#endif
cudaError_t launch_reproject(const KParams& P, cudaStream_t st) {
    const int rows = P.row1 - P.row0;
    ReprojectMaps M;
    if (BMFR_REPROJECT_TMA && (P.W & 3) == 0 && rows >= 32 && P.W >= 32 &&
        bmfr_tensor_map_2d(P.cur_positions, 4, (long long)P.W * 3, rows, RP_TILE_W, 32, &M.positions) &&
        bmfr_tensor_map_2d(P.cur_normals, 4, (long long)P.W * 3, rows, RP_TILE_W, 32, &M.normals) &&
        bmfr_tensor_map_2d(P.cur_noisy, 4, (long long)P.W * 3, rows, RP_TILE_W, 32, &M.noisy)) {
        const dim3 grid((P.W + 31) / 32, (P.k1_y1 - P.k1_y0 + 31) / 32);
        if (is_strip(P)) return launch_pdl(!P.plain_launch, reproject_tma_kernel<true>, grid, dim3(256), 0, st, P, M);
        return launch_pdl(!P.plain_launch, reproject_tma_kernel<false>, grid, dim3(256), 0, st, P, M);
    }
    const int bx = is_strip(P) ? BMFR_REPROJECT_STRIP_BX : BMFR_REPROJECT_BX, by = 256 / bx, rows_per_cta = by * BMFR_REPROJECT_PIXELS;
    const dim3 grid((P.W + bx - 1) / bx, (P.k1_y1 - P.k1_y0 + rows_per_cta - 1) / rows_per_cta), block(bx, by);
    if (is_strip(P)) return launch_pdl(!P.plain_launch, reproject_kernel<true>, grid, block, 0, st, P);
    return launch_pdl(!P.plain_launch, reproject_kernel<false>, grid, block, 0, st, P);
}
// ---- tensor maps of a frame: [rows][W*3] floats, box = one 32x32-pixel tile (bmfr_tma.cuh) ----------------
static bool tile_map(const float* base, int W, int rows, CUtensorMap* out) {
    if ((W & 3) != 0 || rows < 32) return false;
    return bmfr_tensor_map_2d(base, 4, (long long)W * 3, rows, QR_TILE_W, 32, out);
}

// Persistent grid of a fit kernel: as many CTAs as stay resident (never more than blocks), configured once per device.
template <class K0, class K1>
static cudaError_t fit_grid(K0 k_plain, K1 k_strip, int smem, int max_per_sm, int* sm_counts, int* grid_out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (sm_counts[dev] == 0) {
        int n = 0;
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_strip, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        // sized from what the device really keeps resident, not from the launch bounds
        int per_sm = 0;
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_strip, QR_THREADS, (size_t)smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        if (per_sm > max_per_sm) per_sm = max_per_sm;
        sm_counts[dev] = n * per_sm;
    }
    *grid_out = sm_counts[dev];
    return cudaSuccess;
}

template <int FS>
static cudaError_t launch_fit_gram(const KParams& P, const QrMaps& M, int nblocks, cudaStream_t st) {
    static int counts[64] = {};  // resident CTAs per device; 0 = not configured yet
    const int smem = (int)sizeof(GramShared);
    int grid = 0;
    cudaError_t e = fit_grid(fit_gram_kernel<false, FS>, fit_gram_kernel<true, FS>, smem, BMFR_GRAM_MIN_BLOCKS, counts, &grid);
    if (e != cudaSuccess) return e;
    if (grid > nblocks) grid = nblocks;
    if (is_strip(P)) return launch_pdl(!P.plain_launch, fit_gram_kernel<true, FS>, dim3(grid), dim3(QR_THREADS), (size_t)smem, st, P, M);
    return launch_pdl(!P.plain_launch, fit_gram_kernel<false, FS>, dim3(grid), dim3(QR_THREADS), (size_t)smem, st, P, M);
}

cudaError_t launch_fit_qr(const KParams& P, cudaStream_t st) {
    const int nblocks = P.blocks_x * (P.by1 - P.by0);
    if (nblocks < 1) return cudaSuccess;
    QrMaps M;
    memset(&M, 0, sizeof(M));
    const int rows = P.row1 - P.row0;
    M.use_tma = tile_map(P.cur_normals, P.W, rows, &M.normals) && tile_map(P.cur_positions, P.W, rows, &M.positions) &&
                tile_map(P.cur_noisy_acc, P.W, rows, &M.colour);
    if (P.fit_method == BMFR_FIT_GRAM) {
        switch (P.feature_set) {
            case BMFR_FEATURE_SET_LINEAR: return launch_fit_gram<BMFR_FEATURE_SET_LINEAR>(P, M, nblocks, st);
            case BMFR_FEATURE_SET_POSITION: return launch_fit_gram<BMFR_FEATURE_SET_POSITION>(P, M, nblocks, st);
            default: return launch_fit_gram<BMFR_FEATURE_SET_DEFAULT>(P, M, nblocks, st);
        }
    }
    static int counts_qr[64] = {};
    const int smem = (int)sizeof(QrShared);
    int grid = 0;
    cudaError_t e = fit_grid(fit_qr_kernel<false>, fit_qr_kernel<true>, smem, BMFR_QR_MIN_BLOCKS, counts_qr, &grid);
    if (e != cudaSuccess) return e;
    if (grid > nblocks) grid = nblocks;
    if (is_strip(P)) return launch_pdl(!P.plain_launch, fit_qr_kernel<true>, dim3(grid), dim3(QR_THREADS), (size_t)smem, st, P, M);
    return launch_pdl(!P.plain_launch, fit_qr_kernel<false>, dim3(grid), dim3(QR_THREADS), (size_t)smem, st, P, M);
}
