// FUSED mode, second half of the frame: weighted_sum + accumulate_filtered_data + taa
// (bmfr.cl:703-758, 761-857, 860-974) in one pass; `filtered` and `tone_mapped` never reach HBM.
//
// One CTA per 32x32 tile of the frame's shifted block grid, so the block's 42 fit coefficients are
// CTA-uniform: they sit in shared memory (with those of the eight neighbouring blocks, for the ring) and
// are read as warp-wide broadcasts, one read serving a pair of pixels.  Thread (lane, warp) owns the
// column strip x = x0 + lane, rows y0 + 4*warp .. +3: every global access of a warp is 32 consecutive
// pixels of one row (the interleaved-RGB stride of 12 B keeps each 32-bit load on 3-4 cache lines), and
// the four 3x3 TAA neighbourhoods of a strip share their row minima / maxima.
//   phase A : filtered -> accumulated -> tone-mapped for the tile and a one-pixel ring (ring pixels
//             use their own block's coefficients), written to shared memory as YCoCg planes; the TAA
//             history gather of a pixel is issued here too, next to the accumulation gather (both
//             depend only on the stored previous-pixel position: one dependent gather round per
//             pixel).  Neighbours outside the image are filled with the nearest in-image pixel, which
//             leaves the min / max over the in-image neighbours unchanged (bmfr.cl:900-920) and
//             removes every per-neighbour test.
//   phase B : clamp of the history sample to the neighbourhood box from shared memory, blend, store.
// Nothing here is compared bitwise with the reference (the inputs already carry the fit's
// rounding), so this translation unit is compiled with FMA contraction and uses the fast
// reciprocal / lg2 / ex2; tests/test_gpu_parity.py holds it to the 1e-3 / 60 dB colour tolerance.
#include "bmfr_kernels.h"

#include "bmfr_device.cuh"
#include "bmfr_tma.cuh"

#define PT_TILE 32
#define PT_HALO (PT_TILE + 2)
#define PT_STRIDE 36   // floats per shared-memory row (>= 34)
#define PT_COEF 52     // per block: F x (w_r, w_g, w_b, -) then NSC x (min, 1/range); 10 and 6 for the largest list

struct PostShared {
    float ycc[3][PT_HALO][PT_STRIDE];
    float coef[9][PT_COEF];  // this block ([4]) and its eight neighbours (ring pixels), (dy+1)*3 + dx+1
};

// WIDE: every interleaved-RGB buffer of the launch is 8-byte aligned, so a pixel is moved with a
// 64-bit + a 32-bit access (bmfr_device.cuh) instead of three 32-bit ones.
template <bool WIDE>
__device__ __forceinline__ f3 ldf3(const float* __restrict__ b, unsigned int i) {
    return WIDE ? load_f3_wide(b, i) : load_f3(b, i);
}

template <bool WIDE>
__device__ __forceinline__ void stf3(float* __restrict__ b, unsigned int i, f3 v) {
    if (WIDE) store_f3_wide(b, i, v);
    else store_f3(b, i, v);
}

// Strips: a row a neighbour mirrors is stored a second time, into the neighbour's halo (peer memory over NVLink).
// which = 0: accumulated filtered colour, 1: TAA result.
template <bool STRIP>
__device__ __forceinline__ void post_push(const KParams& P, int which, int x, int y, f3 v) {
    if (!STRIP || !P.halo_p.active) return;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const long long pi = halo_peer_index(P.halo_p, P, s, x, y);
        if (pi >= 0) {
            float* dst = (which ? P.halo_p.peer_b[s] : P.halo_p.peer_a[s]) + pi * 3;
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z;
        }
    }
}

__device__ __forceinline__ float fast_rcp(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ float tone_map_fast(float v) {  // clamp(powr(max(0,v), 0.454545f), 0, 1), bmfr.cl:852-856
    v = fmaxf(0.f, v);
    float l, e;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(v));  // 0 -> -inf -> ex2 -> 0, like powr(0, y > 0)
    l *= 0.454545f;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(l));
    return __saturatef(e);
}
// The block's coefficients in shared memory (PostShared::coef / PostStage::coef), for a list of F features of which the
// last NSC are scaled: F x (w_r, w_g, w_b, -), then NSC x (min, 1/range) — read as 128-bit broadcasts (every lane of a
// tile-interior warp reads the same address).
template <int FS>
struct PostCoef {
    static constexpr int F = FeatureSet<FS>::F, NSC = FeatureSet<FS>::NSC, NW = 3 * F, NM = 2 * NSC, MINV = 4 * F;
    float4 m[(NSC + 1) / 2];  // (min, 1/range) pairs, two per float4
    __device__ __forceinline__ void load(const float* __restrict__ cf) {
        const float4* c4 = reinterpret_cast<const float4*>(cf);
#pragma unroll
        for (int j = 0; j < (NSC + 1) / 2; ++j) m[j] = c4[F + j];
    }
    __device__ __forceinline__ float mn(int k) const { return (k & 1) ? m[k >> 1].z : m[k >> 1].x; }
    __device__ __forceinline__ float inv(int k) const { return (k & 1) ? m[k >> 1].w : m[k >> 1].y; }
    // the F - 1 non-constant features of a pixel (bmfr.cl:724-741): clean, no noise, no NaN scrub
    __device__ __forceinline__ void features(f3 n, f3 p, float (&feat)[F - 1]) const {
        int at = 0;
        if (FeatureSet<FS>::NORMALS) { feat[0] = n.x; feat[1] = n.y; feat[2] = n.z; at = 3; }
        const float raw[6] = {p.x, p.y, p.z, p.x * p.x, p.y * p.y, p.z * p.z};
#pragma unroll
        for (int k = 0; k < NSC; ++k) feat[at + k] = (raw[k] - mn(k)) * inv(k);
    }
};
__device__ __forceinline__ f3 clamp_negative(f3 c) {  // keeps NaN like the reference, bmfr.cl:750
    return make_f3(c.x < 0.f ? 0.f : c.x, c.y < 0.f ? 0.f : c.y, c.z < 0.f ? 0.f : c.z);
}

// weighted_sum for one pixel, bmfr.cl:725-750.
template <int FS>
__device__ __forceinline__ f3 weighted_sum_px(f3 n, f3 p, const float* __restrict__ cf) {
    constexpr int F = FeatureSet<FS>::F;
    const float4* c4 = reinterpret_cast<const float4*>(cf);
    PostCoef<FS> pc;
    pc.load(cf);
    float feat[F - 1];
    pc.features(n, p, feat);
    const float4 w0 = c4[0];
    f3 c = make_f3(w0.x, w0.y, w0.z);
#pragma unroll
    for (int f = 1; f < F; ++f) {
        const float4 w = c4[f];
        c.x = fmaf(w.x, feat[f - 1], c.x);
        c.y = fmaf(w.y, feat[f - 1], c.y);
        c.z = fmaf(w.z, feat[f - 1], c.z);
    }
    return clamp_negative(c);
}

// The same for two pixels of a strip: every coefficient is fetched once and used twice, which halves
// the shared-memory wavefronts of the weighted sum (a warp-wide 128-bit broadcast is four wavefronts).
template <int FS>
__device__ __forceinline__ void weighted_sum_px2(f3 n0, f3 p0, f3 n1, f3 p1, const float* __restrict__ cf, f3& out0, f3& out1) {
    constexpr int F = FeatureSet<FS>::F;
    const float4* c4 = reinterpret_cast<const float4*>(cf);
    PostCoef<FS> pc;
    pc.load(cf);
    float f0[F - 1], f1[F - 1];
    pc.features(n0, p0, f0);
    pc.features(n1, p1, f1);
    const float4 w0 = c4[0];
    f3 a = make_f3(w0.x, w0.y, w0.z), b = a;
#pragma unroll
    for (int f = 1; f < F; ++f) {
        const float4 w = c4[f];
        a.x = fmaf(w.x, f0[f - 1], a.x); a.y = fmaf(w.y, f0[f - 1], a.y); a.z = fmaf(w.z, f0[f - 1], a.z);
        b.x = fmaf(w.x, f1[f - 1], b.x); b.y = fmaf(w.y, f1[f - 1], b.y); b.z = fmaf(w.z, f1[f - 1], b.z);
    }
    out0 = clamp_negative(a);
    out1 = clamp_negative(b);
}

// The coefficients of the 3x3 block neighbourhood -> shared memory: warp w takes neighbour w (warp 0 also the ninth).
template <int FS>
__device__ __forceinline__ void load_coefficients(const KParams& P, float (*coef)[PT_COEF], int bx, int by, int warp, int lane, int nwarps = 8) {
    using PC = PostCoef<FS>;
    for (int nb = warp; nb < 9; nb += nwarps) {
        const int gx = bx + nb % 3 - 1, gy = by + nb / 3 - 1;
        if (gx < 0 || gx >= P.blocks_x || gy < 0 || gy >= P.blocks_y) continue;
        const size_t g = (size_t)gy * P.blocks_x + gx;
        if (lane < PC::NW) coef[nb][(lane / 3) * 4 + lane % 3] = __ldg(P.weights + g * PC::NW + lane);
        if (lane < PC::NM) coef[nb][PC::MINV + lane] = __ldg(P.mins_inv + g * PC::NM + lane);
    }
}

// accept / pp / spp / alb are this pixel's accept mask, previous-frame position, sample count and
// albedo, fetched by the caller together with the features (one round of independent loads).
template <bool STRIP, bool WIDE>
__device__ __forceinline__ f3 accumulate_filtered_px(const KParams& P, unsigned int lp, f3 filtered, unsigned int accept,
                                                     float2 pp, unsigned int spp, f3 alb, bool store, int x, int y,
                                                     bool push = true, f3* accum_out = nullptr) {
    f3 prev = make_f3(0.f, 0.f, 0.f);
    float alpha = 1.f;
    if (P.frame > 0 && accept != 0) {
        const int pix = __float2int_rd(pp.x), piy = __float2int_rd(pp.y);
        const float frx = pp.x - (float)pix, fry = pp.y - (float)piy;
        const float omx = 1.f - frx, omy = 1.f - fry;
        const float w[4] = {omx * omy, frx * omy, omx * fry, frx * fry};
        float total = 0.f;
        // (a strip that does not hold a tap's row reports it once, after the loop: a store inside would keep the compiler
        // from predicating the taps)
        bool missing = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int sx = pix + (i & 1), sy = piy + (i >> 1);
            const bool held = !STRIP || (sy >= P.state2_row0 && sy < P.state2_row1);
            const bool want = (accept & (1u << i)) != 0;  // taps are trusted, not re-checked: bmfr.cl:801-832
            missing = missing || (want && !held);
            if (want && held) {
                const f3 pc = ldf3<WIDE>(P.accum_prev, pix_index(P, sx, sy));
                total += w[i];
                prev.x = fmaf(w[i], pc.x, prev.x);
                prev.y = fmaf(w[i], pc.y, prev.y);
                prev.z = fmaf(w[i], pc.z, prev.z);
            }
        }
        if (STRIP && missing) *P.oob_flag = 1;
        if (total > 0.f) {
            alpha = fmaxf(fast_rcp((float)spp), P.second_blend_alpha);  // bmfr.cl:838-839
            const float inv = fast_rcp(total);
            prev.x *= inv;
            prev.y *= inv;
            prev.z *= inv;
        }
    }
    const float oma = 1.f - alpha;
    const f3 accum = make_f3(fmaf(alpha, filtered.x, oma * prev.x), fmaf(alpha, filtered.y, oma * prev.y),
                             fmaf(alpha, filtered.z, oma * prev.z));
    if (store) {
        stf3<WIDE>(P.accum_cur, lp, accum);
        if (push) post_push<STRIP>(P, 0, x, y, accum);
    }
    if (accum_out) *accum_out = accum;
    return make_f3(tone_map_fast(alb.x * accum.x), tone_map_fast(alb.y * accum.y), tone_map_fast(alb.z * accum.z));
}

__device__ __forceinline__ f3 to_ycocg(f3 c) {  // bmfr.cl:184-190: (r + 2g + b, 2r - 2b, -r + 2g - b), five operations
    const float t = c.x + c.z, d = c.x - c.z;
    return make_f3(fmaf(2.f, c.y, t), d + d, fmaf(2.f, c.y, -t));
}
__device__ __forceinline__ f3 from_ycocg(f3 c) {  // bmfr.cl:192-198
    return make_f3(0.25f * (c.x + c.y - c.z), 0.25f * (c.x + c.z), 0.25f * (c.x - c.y - c.z));
}

__device__ __forceinline__ void put_cell(PostShared& sh, int hx, int hy, f3 v) {
    sh.ycc[0][hy][hx] = v.x;
    sh.ycc[1][hy][hx] = v.y;
    sh.ycc[2][hy][hx] = v.z;
}
// Stores the YCoCg value of image pixel (x,y) at halo cell (hx,hy) and replicates it into the
// out-of-image cells whose nearest in-image pixel it is.
__device__ __forceinline__ void put_ycc(PostShared& sh, const KParams& P, int hx, int hy, int x, int y, f3 v) {
    put_cell(sh, hx, hy, v);
    const int ex = (x == 0) ? -1 : (x == P.W - 1) ? 1 : 0;
    const int ey = (y == 0) ? -1 : (y == P.H - 1) ? 1 : 0;
    if ((ex | ey) == 0) return;
    const bool okx = ex != 0 && (unsigned)(hx + ex) < PT_HALO, oky = ey != 0 && (unsigned)(hy + ey) < PT_HALO;
    if (okx) put_cell(sh, hx + ex, hy, v);
    if (oky) put_cell(sh, hx, hy + ey, v);
    if (okx && oky) put_cell(sh, hx + ex, hy + ey, v);
}

// Bilinear sample of the previous TAA result at pp in YCoCg (bmfr.cl:922-965); false when the pixel
// takes the copy-through path of bmfr.cl:884-890.  It depends on pp only, so it is issued in phase A
// next to the accumulation taps: one round of gathers per pixel instead of two.
template <bool STRIP, bool WIDE>
__device__ __forceinline__ bool history_sample(const KParams& P, float2 pp, f3& hist) {
    hist = make_f3(0.f, 0.f, 0.f);
    if (P.frame == 0) return false;
    const int pix = __float2int_rd(pp.x), piy = __float2int_rd(pp.y);
    if (pix < -1 || piy < -1 || pix >= P.W || piy >= P.H) return false;  // bmfr.cl:884-890
    const float frx = pp.x - (float)pix, fry = pp.y - (float)piy;
    const float omx = 1.f - frx, omy = 1.f - fry;
    const float w[4] = {omx * omy, frx * omy, omx * fry, frx * fry};
    f3 prev = make_f3(0.f, 0.f, 0.f);
    float total = 0.f;
    bool missing = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bmfr.cl:929-960
        const int dx = i & 1, dy = i >> 1;
        const bool ok_y = dy ? (piy < P.H - 1) : (piy >= 0);
        const bool ok_x = dx ? (pix < P.W - 1) : (pix >= 0);
        const int sy = piy + dy;
        const bool held = !STRIP || (sy >= P.state2_row0 && sy < P.state2_row1);
        missing = missing || (ok_x && ok_y && !held);
        if (ok_x && ok_y && held) {
            const f3 pc = ldf3<WIDE>(P.result_prev, pix_index(P, pix + dx, sy));
            prev.x = fmaf(w[i], pc.x, prev.x);
            prev.y = fmaf(w[i], pc.y, prev.y);
            prev.z = fmaf(w[i], pc.z, prev.z);
            total += w[i];
        }
    }
    if (STRIP && missing) *P.oob_flag = 1;
    const float inv = fast_rcp(total);  // 0 * inf = NaN on the image edge like the 0/0 of bmfr.cl:962
    hist = to_ycocg(make_f3(prev.x * inv, prev.y * inv, prev.z * inv));
    return true;
}

// weighted_sum -> accumulate_filtered_data -> tone map of image pixel (x,y) into halo cell (hx,hy);
// with `own` also the pixel's TAA history sample (returns whether it takes the temporal path)
struct PixelIn {  // everything phase A reads at the pixel itself: one round of independent loads
    f3 n, p, alb;
    float2 pp;
    unsigned int lp, accept, spp;
};
template <bool WIDE>
__device__ __forceinline__ PixelIn load_pixel(const KParams& P, int x, int y) {
    PixelIn in;
    in.lp = pix_index(P, x, y);
    in.n = WIDE ? load_f3_wide(P.cur_normals, in.lp) : load_f3_stream(P.cur_normals, in.lp);
    in.p = WIDE ? load_f3_wide(P.cur_positions, in.lp) : load_f3_stream(P.cur_positions, in.lp);
    in.alb = WIDE ? load_f3_wide(P.albedo, in.lp) : load_f3_stream(P.albedo, in.lp);
    in.accept = __ldg(P.accept + in.lp);
    in.spp = __ldg(const_cast<const unsigned char*>(P.cur_spp) + in.lp);
    in.pp = __ldg(P.prev_pixels + in.lp);
    return in;
}
template <bool STRIP, bool WIDE>
__device__ __forceinline__ bool finish_pixel(PostShared& sh, const KParams& P, const PixelIn& in, f3 filtered, int hx, int hy, int x,
                                             int y, bool store, bool own, f3& hist) {
    const bool temporal = own && history_sample<STRIP, WIDE>(P, in.pp, hist);
    const f3 tone = accumulate_filtered_px<STRIP, WIDE>(P, in.lp, filtered, in.accept, in.pp, in.spp, in.alb, store, x, y);
    put_ycc(sh, P, hx, hy, x, y, to_ycocg(tone));
    return temporal;
}
template <bool STRIP, bool WIDE, int FS>
__device__ __forceinline__ bool phase_a_pixel(PostShared& sh, const KParams& P, const float* cf, int hx, int hy, int x, int y,
                                              bool store, bool own, f3& hist) {
    const PixelIn in = load_pixel<WIDE>(P, x, y);
    return finish_pixel<STRIP, WIDE>(sh, P, in, weighted_sum_px<FS>(in.n, in.p, cf), hx, hy, x, y, store, own, hist);
}

#ifndef BMFR_POST_PREFETCH_TAPS
#define BMFR_POST_PREFETCH_TAPS 0  // measured: no effect
#endif
#ifndef BMFR_POST_PREFETCH
#define BMFR_POST_PREFETCH 1  // L2 prefetch of the pixel inputs before the wait for the fit (see the kernel)
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// Starts the DRAM -> L2 fetch of everything phase A reads at pixel (x,y) (no registers are tied up); the
// loads that follow one or two pixels later then see L2 latency instead of DRAM latency.
__device__ __forceinline__ void prefetch_pixel(const KParams& P, int x, int y) {
    const unsigned int lp = pix_index(P, x, y);
    prefetch_l2(P.cur_normals + (size_t)(lp * 3u));
    prefetch_l2(P.cur_positions + (size_t)(lp * 3u));
    prefetch_l2(P.albedo + (size_t)(lp * 3u));
    prefetch_l2(P.prev_pixels + lp);
}
// The same for the two gathers of a pixel whose previous-frame position is already known.
__device__ __forceinline__ void prefetch_taps(const KParams& P, float2 pp) {
    const int pix = min(max(__float2int_rd(pp.x), 0), P.W - 1);
    const int piy = __float2int_rd(pp.y);
    const int ya = min(max(piy, P.row0), P.row1 - 1), yb = min(max(piy + 1, P.row0), P.row1 - 1);
    const unsigned int la = pix_index(P, pix, ya), lb = pix_index(P, pix, yb);
    prefetch_l2(P.accum_prev + (size_t)(la * 3u));
    prefetch_l2(P.accum_prev + (size_t)(lb * 3u));
    prefetch_l2(P.result_prev + (size_t)(la * 3u));
    prefetch_l2(P.result_prev + (size_t)(lb * 3u));
}

#ifndef BMFR_POST_WIDE_ACCESS
#define BMFR_POST_WIDE_ACCESS 0
#endif
#ifndef BMFR_POST_MIN_BLOCKS
#define BMFR_POST_MIN_BLOCKS 5
#endif

template <bool STRIP, bool WIDE, int FS>
__global__ void __launch_bounds__(256, BMFR_POST_MIN_BLOCKS) post_kernel(const __grid_constant__ KParams P) {
    __shared__ __align__(16) PostShared sh;
    const int bx = blockIdx.x, by = P.by0 + (STRIP ? halo_row_order(P.halo_p, blockIdx.y, gridDim.y) : sweep_row(P, blockIdx.y, gridDim.y));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = bx * 32 - 16 + P.off_x, y0 = by * 32 - 16 + P.off_y;  // tile origin in image coordinates
#if BMFR_POST_PREFETCH
    // DRAM -> L2 prefetch of the strip's four pixels and of this thread's ring pixel, issued before the
    // wait for the fit: CTAs that became resident while the fit drains warm the L2 for their pixels
    {
        const int px = x0 + lane;
        if (px >= 0 && px < P.W) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int py = y0 + 4 * warp + s;
                if (py >= P.py0 && py < P.py1) prefetch_pixel(P, px, py);
            }
        }
    }
#endif
    pdl_wait();     // the fit of this frame is complete (weights, min/max)
    // Only now: a successor (the next frame's reprojection) starts once every CTA of this grid has passed
    // this point, so "this frame's fit and reprojection are complete" holds for it too.
    pdl_trigger();
    stamp_begin(P, 2);
    // strips: a CTA near a strip edge waits for the neighbours' accumulated colour / TAA rows of the previous frame
    const bool zone = STRIP && halo_in_zone(P.halo_p, y0 - 1, y0 + 33);
    if (zone) halo_poll(P.halo_p, halo_peek(P.halo_p, zone));


    load_coefficients<FS>(P, sh.coef, bx, by, warp, lane);
    __syncthreads();

    const int x = x0 + lane;
    const bool col_ok = x >= 0 && x < P.W;
    // phase A, interior: column strip x, rows 4*warp .. 4*warp+3
    f3 hist[4];
    unsigned int live = 0;  // bit s: pixel s is written; bit 4+s: it takes the temporal path
#pragma unroll
    for (int s = 0; s < 4; s += 2) {  // two pixels at a time: they share the coefficient loads
        const int ty = 4 * warp + s, y = y0 + ty;
        hist[s] = hist[s + 1] = make_f3(0.f, 0.f, 0.f);
        const bool v0 = col_ok && y >= P.py0 && y < P.py1, v1 = col_ok && y + 1 >= P.py0 && y + 1 < P.py1;
        if (v0 && v1) {
            const PixelIn i0 = load_pixel<WIDE>(P, x, y), i1 = load_pixel<WIDE>(P, x, y + 1);
#if BMFR_POST_PREFETCH_TAPS
            if (s == 0 && P.frame > 0) {  // while this pair is processed, pull the next pair's gather targets into L2
                const int yn = y + 2;
                if (yn >= P.py0 && yn + 1 < P.py1) {
                    prefetch_taps(P, __ldg(P.prev_pixels + pix_index(P, x, yn)));
                    prefetch_taps(P, __ldg(P.prev_pixels + pix_index(P, x, yn + 1)));
                }
            }
#endif
            f3 fl0, fl1;
            weighted_sum_px2<FS>(i0.n, i0.p, i1.n, i1.p, sh.coef[4], fl0, fl1);
            const bool own0 = y >= P.own_y0 && y < P.own_y1, own1 = y + 1 >= P.own_y0 && y + 1 < P.own_y1;
            const bool t0 = finish_pixel<STRIP, WIDE>(sh, P, i0, fl0, lane + 1, ty + 1, x, y, true, own0, hist[s]);
            const bool t1 = finish_pixel<STRIP, WIDE>(sh, P, i1, fl1, lane + 1, ty + 2, x, y + 1, true, own1, hist[s + 1]);
            live |= ((own0 ? 1u : 0u) | (t0 ? 16u : 0u)) << s;
            live |= ((own1 ? 1u : 0u) | (t1 ? 16u : 0u)) << (s + 1);
        } else if (v0) {  // a strip or image edge cuts the pair
            const bool own = y >= P.own_y0 && y < P.own_y1;
            const bool t = phase_a_pixel<STRIP, WIDE, FS>(sh, P, sh.coef[4], lane + 1, ty + 1, x, y, true, own, hist[s]);
            live |= ((own ? 1u : 0u) | (t ? 16u : 0u)) << s;
        } else if (v1) {
            const bool own = y + 1 >= P.own_y0 && y + 1 < P.own_y1;
            const bool t = phase_a_pixel<STRIP, WIDE, FS>(sh, P, sh.coef[4], lane + 1, ty + 2, x, y + 1, true, own, hist[s + 1]);
            live |= ((own ? 1u : 0u) | (t ? 16u : 0u)) << (s + 1);
        }
    }
    // phase A, ring: 4 * 33 = 132 pixels, coefficients of the pixel's own block
    if (tid < 4 * (PT_HALO - 1)) {
        const int side = tid / (PT_HALO - 1), k = tid % (PT_HALO - 1);
        int hx, hy;
        if (side == 0) { hx = k; hy = 0; }
        else if (side == 1) { hx = PT_HALO - 1; hy = k; }
        else if (side == 2) { hx = PT_HALO - 1 - k; hy = PT_HALO - 1; }
        else { hx = 0; hy = PT_HALO - 1 - k; }
        const int rx = x0 + hx - 1, ry = y0 + hy - 1;
        if (rx >= 0 && rx < P.W && ry >= P.py0 && ry < P.py1) {
            const int nb = ((hy == 0) ? 0 : (hy == PT_HALO - 1) ? 6 : 3) + ((hx == 0) ? 0 : (hx == PT_HALO - 1) ? 2 : 1);
            f3 unused;
            phase_a_pixel<STRIP, WIDE, FS>(sh, P, sh.coef[nb], hx, hy, rx, ry, false, false, unused);
        }
    }
    __syncthreads();

    // phase B: clamp the history samples to the neighbourhood box, plane by plane (bmfr.cl:893-920, 967-969).  Halo rows
    // 4*warp .. 4*warp+5 cover the 3x3 neighbourhoods of the strip; this thread's column is lane+1.
    f3 mine[4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float ctr[6], rmin[6], rmax[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const float* row = sh.ycc[c][4 * warp + r];
            const float l = row[lane], m = row[lane + 1], rr = row[lane + 2];
            ctr[r] = m;
            rmin[r] = fminf(fminf(l, m), rr);
            rmax[r] = fmaxf(fmaxf(l, m), rr);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const float min_box = fminf(fminf(rmin[s], rmin[s + 1]), rmin[s + 2]);
            const float max_box = fmaxf(fmaxf(rmax[s], rmax[s + 1]), rmax[s + 2]);
            const float min_cross = fminf(fminf(ctr[s], rmin[s + 1]), ctr[s + 2]);
            const float max_cross = fmaxf(fmaxf(ctr[s], rmax[s + 1]), ctr[s + 2]);
            const float lo = (min_box + min_cross) * 0.5f, hi = (max_box + max_cross) * 0.5f;
            float& h = (c == 0) ? hist[s].x : (c == 1) ? hist[s].y : hist[s].z;
            h = fminf(fmaxf(h, lo), hi);
            float& m = (c == 0) ? mine[s].x : (c == 1) ? mine[s].y : mine[s].z;
            m = ctr[s + 1];
        }
    }
    // step 3: blend and store (bmfr.cl:971-973)
    const float a = P.taa_blend_alpha, oma = 1.f - P.taa_blend_alpha;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (!(live & (1u << s))) continue;
        const unsigned int lp = pix_index(P, x, y0 + 4 * warp + s);
        f3 out = from_ycocg(mine[s]);  // this pixel's tone-mapped colour
        if (live & (16u << s)) {
            const f3 pr = from_ycocg(hist[s]);
            out = make_f3(fmaf(a, out.x, oma * pr.x), fmaf(a, out.y, oma * pr.y), fmaf(a, out.z, oma * pr.z));
        }
        stf3<WIDE>(P.result_cur, lp, out);
        if (P.user_out) stf3<WIDE>(P.user_out, lp, out);
        post_push<STRIP>(P, 1, x, y0 + 4 * warp + s, out);
    }
    if (zone) halo_finish(P.halo_p, halo_cta_pushes(P.halo_p, y0, y0 + 32));
    stamp_end(P, 2);
}


// ================================================================================================
// TMA-staged variant (the default wherever the tensor maps can be built: W % 16 == 0, at least 34 rows).
//
// The read-once inputs of a tile — normals, positions, albedo, previous-pixel positions, accept masks and sample
// counts of the 34x34 pixels around it (tile + ring) — arrive as six bulk tensor copies (cp.async.bulk.tensor,
// one elected thread, one mbarrier) instead of ~2000 L1 wavefronts of per-thread loads: an interleaved-RGB
// image read component by component touches 3-4 cache lines per 32-bit warp load, and the ring columns one line
// per pixel.  The copies are requested before the wait for the fit (everything they read is complete by then,
// see the kernel), so a tile's inputs land while its neighbours on the SM compute.  From shared memory a warp reads
// the same values at a stride of three words, which is bank-conflict free.
//   * Each pixel's tone-mapped YCoCg value is written over its own albedo cell (only that pixel's thread ever
//     reads the cell), so the neighbourhood planes of phase B need no memory of their own.
//   * What still goes through the L1 are the two 4-tap gathers.  A thread's vertically adjacent pixels share
//     their middle tap row whenever the reprojection is locally uniform (the common case; checked per pair), so
//     a pair loads three tap rows instead of four; all taps come in one round of independent loads, issued
//     before the weighted sum so that its arithmetic overlaps their latency.
//   * Interior tiles (BMFR_POST_FAST) — halo inside the image, every pixel owned, no strip edge near: 88 % of the tiles
//     at 1080p — take a phase A without the per-pixel image-edge / strip / ownership tests of the general one: the
//     source-level profile of the general body (profiles/r02_final_ncu_full.md, 486 warp instructions per pixel) has a
//     quarter of its issue slots in predicates, clamps, index arithmetic, the replication of edge cells and the
//     divergence bookkeeping of rarely taken branches.  There, after a warp-uniform test per pixel pair (every lane's
//     stacked 2x3 footprint inside the rows / columns held) the 36 tap loads use three row pointers per buffer and
//     immediate offsets, every tap of the TAA history is valid and nothing is clamped; pairs that fail the test take
//     the general pair code.
//   * ONE arithmetic per pixel.  Whichever path computes a pixel — interior or general tile, pair or single, tile or
//     ring, whole image or strip — it runs the same rounded operations in the same order (the helpers below, written
//     with explicit intrinsics so that the compiler's contraction cannot differ between inlining sites): a strip run
//     stays bit-identical to the whole-image run although the two sort different tiles into the two paths
//     (tests/test_sharding.py, tests/test_baseline_configs.py).
//     The block's coefficients are pre-multiplied by 1/range when they are loaded (one record of F x (w_r, w_g, w_b, min)
//     per block: a scaled feature is one subtraction), the accumulation is written without branches (selects on
//     total_weight > 0 instead of the nested ifs of bmfr.cl:784-841), the clamp of the tone map sits in front of the
//     power (a saturating multiply), a TAA history sample whose four taps are all valid is not divided by its weight
//     sum (1 +- 2 ulp), and the TAA blend is done in YCoCg (the transform is linear: one conversion back instead of two).
// Everything here is tolerance-only arithmetic (tests/test_gpu_parity.py: 1e-3 / 60 dB against the oracle).
// ================================================================================================
#define PT_RGB_W 108  // floats per staged row of an interleaved-RGB image: 34 pixels + up to 3 floats of alignment shift
#define PT_PP_W 72    // floats per staged row of prev_pixels: 34 float2 + 2 floats of shift
#define PT_U8_W 64    // bytes per staged row of accept / spp: 34 + up to 15 bytes of shift
#define PT_COEF_S 40  // floats of a block's pre-scaled coefficient record: F x (w_r, w_g, w_b, min), F <= 10

// A tile is 32 x 32 pixels; the stage holds it with its one-pixel ring: HY = 34 rows.
template <int HY_>
struct __align__(128) PostStageT {  // every TMA destination starts on a 128-byte boundary
    static constexpr int HY = HY_;
    static constexpr int TX = 3 * HY_ * PT_RGB_W * 4 + HY_ * PT_PP_W * 4 + 2 * HY_ * PT_U8_W;  // bytes of the six bulk copies
    float nrm[HY_][PT_RGB_W]; char pad0[32];
    float pos[HY_][PT_RGB_W]; char pad1[32];
    float alb[HY_][PT_RGB_W]; char pad2[32];  // albedo, then (cell by cell) the tone-mapped colour as YCoCg
    float pp[HY_][PT_PP_W];   char pad3[64];
    unsigned char acc[HY_][PT_U8_W];
    unsigned char spp[HY_][PT_U8_W];
    float coef[2][9][PT_COEF_S];  // two tiles' worth: the persistent kernel loads the next tile's while this one is in phase B
    unsigned long long bar;      // the tile's inputs have landed (persistent kernel: all but the albedo)
    unsigned long long bar_alb;  // persistent kernel: the tile's albedo has landed
    static_assert(sizeof(float[HY_][PT_RGB_W]) % 128 == 96 && sizeof(float[HY_][PT_PP_W]) % 128 == 64 && (HY_ * PT_U8_W) % 128 == 0,
                  "TMA destinations must stay 128-byte aligned");
};
using PostStage = PostStageT<PT_HALO>;

struct PostMaps {
    CUtensorMap normals, positions, albedo, pp, accept, spp;
};

struct TileGeom {
    int x0, y0;              // image coordinates of the tile's first pixel
    int sh_rgb, sh_pp, sh_u8;  // offset of halo column 0 inside a staged row (alignment of the box start)
};

__device__ __forceinline__ f3 cell_f3(const float (*buf)[PT_RGB_W], const TileGeom& G, int hx, int hy) {
    const float* p = &buf[hy][G.sh_rgb + 3 * hx];
    return make_f3(p[0], p[1], p[2]);
}
__device__ __forceinline__ void put_cell_i(PostStage& sh, const TileGeom& G, int hx, int hy, f3 v) {
    float* p = &sh.alb[hy][G.sh_rgb + 3 * hx];
    p[0] = v.x; p[1] = v.y; p[2] = v.z;
}
// put_ycc() for the interleaved cells: the value of image pixel (x,y) also fills the out-of-image cells whose
// nearest in-image pixel it is (nobody reads an albedo there).
__device__ __forceinline__ void put_ycc_i(PostStage& sh, const KParams& P, const TileGeom& G, int hx, int hy, int x, int y, f3 v) {
    put_cell_i(sh, G, hx, hy, v);
    const int ex = (x == 0) ? -1 : (x == P.W - 1) ? 1 : 0;
    const int ey = (y == 0) ? -1 : (y == P.H - 1) ? 1 : 0;
    if ((ex | ey) == 0) return;
    const bool okx = ex != 0 && (unsigned)(hx + ex) < PT_HALO, oky = ey != 0 && (unsigned)(hy + ey) < PT_HALO;
    if (okx) put_cell_i(sh, G, hx + ex, hy, v);
    if (oky) put_cell_i(sh, G, hx, hy + ey, v);
    if (okx && oky) put_cell_i(sh, G, hx + ex, hy + ey, v);
}
__device__ __forceinline__ PixelIn load_pixel_staged(const PostStage& sh, const KParams& P, const TileGeom& G, int hx, int hy, int x, int y) {
    PixelIn in;
    in.lp = pix_index(P, x, y);
    in.n = cell_f3(sh.nrm, G, hx, hy);
    in.p = cell_f3(sh.pos, G, hx, hy);
    in.pp = *reinterpret_cast<const float2*>(&sh.pp[hy][G.sh_pp + 2 * hx]);
    in.accept = sh.acc[hy][G.sh_u8 + hx];
    in.spp = sh.spp[hy][G.sh_u8 + hx];
    return in;
}
// The pixel's albedo (its cell later holds the pixel's YCoCg value).  The persistent kernel fetches a tile's albedo box behind
// the other inputs (the cells are busy with the previous tile's phase B until then), on a barrier of its own: gate.bar != nullptr.
struct AlbGate {
    unsigned long long* bar;
    unsigned int parity;
};
__device__ __forceinline__ f3 albedo_cell(const PostStage& sh, const TileGeom& G, const AlbGate& gate, int hx, int hy) {
    if (gate.bar != nullptr) mbar_wait_hot(gate.bar, gate.parity);
    return cell_f3(sh.alb, G, hx, hy);
}
// Zone CTAs of a strip (HaloK) do not send a mirrored row pixel by pixel: a pixel's accumulated colour goes into its own
// (already consumed) normal cell, its TAA result into its position cell, and after the last pixel the CTA sends whole row
// segments as 8-byte peer stores (post_push_rows) — 4-byte stores at a 12-byte stride make poor NVLink packets.
__device__ __forceinline__ void stage_accum(PostStage& sh, const TileGeom& G, int hx, int hy, f3 v) {
    float* p = &sh.nrm[hy][G.sh_rgb + 3 * hx];
    p[0] = v.x; p[1] = v.y; p[2] = v.z;
}
__device__ __forceinline__ void stage_result(PostStage& sh, const TileGeom& G, int hx, int hy, f3 v) {
    float* p = &sh.pos[hy][G.sh_rgb + 3 * hx];
    p[0] = v.x; p[1] = v.y; p[2] = v.z;
}
// All threads of a zone CTA, after a barrier: the tile's rows that a neighbour mirrors, both buffers.
template <int THREADS>
__device__ __forceinline__ void post_push_rows(const KParams& P, const PostStage& sh, const TileGeom& G, int tid) {
    constexpr int ROWS = PT_TILE;
    const HaloK& h = P.halo_p;
#ifdef BMFR_DEBUG_NO_PUSH
    return;
#endif
    const int xa = max(G.x0, 0), xb = min(G.x0 + 32, P.W);  // the tile's columns inside the image (both even)
    const int per_row = (xb - xa) * 3 / 2;                  // 8-byte items per row and buffer
    if (per_row <= 0) return;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        if (!h.side_on[s]) continue;
        const int ya = max(h.push_y0[s], G.y0), yb = min(h.push_y1[s], G.y0 + ROWS);
        const int items = (yb - ya) * per_row * 2;
        for (int i = tid; i < items; i += THREADS) {
            const int which = i / ((yb - ya) * per_row), j = i % ((yb - ya) * per_row);
            const int y = ya + j / per_row, k = j % per_row;
            const float* src = (which ? &sh.pos[y - G.y0 + 1][0] : &sh.nrm[y - G.y0 + 1][0]) + G.sh_rgb + 3 * (xa - G.x0 + 1);
            float* dst = (which ? h.peer_b[s] : h.peer_a[s]) + ((long long)(y - h.peer_row0[s]) * P.W + xa) * 3;
            reinterpret_cast<float2*>(dst)[k] = reinterpret_cast<const float2*>(src)[k];
        }
    }
}

// ---- the arithmetic of one pixel (shared by every path; explicit intrinsics, see the banner) ----------------------------
// F x (w_r * s_f, w_g * s_f, w_b * s_f, min_f) per block, s_f = 1/range (scale_factor()) for the scaled features, 1 and
// min = 0 otherwise: warp w takes neighbour w of the 3x3 block neighbourhood, (dy + 1) * 3 + dx + 1 (warp 0 also the ninth).
struct CoefRegs {  // what one lane holds of the (up to two) neighbour blocks its warp loads
    float w[2], s[2], m[2];
};
template <int FS>
__device__ __forceinline__ CoefRegs fetch_coefficients(const KParams& P, int bx, int by, int warp, int lane, int nwarps) {
    constexpr int F = FeatureSet<FS>::F, NSC = FeatureSet<FS>::NSC, NNS = F - NSC;
    const int f = lane / 3;
    CoefRegs r;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int nb = warp + j * nwarps;
        r.w[j] = 0.f; r.s[j] = 1.f; r.m[j] = 0.f;
        if (nb >= 9) continue;
        const int q = (nb * 11) >> 5;  // nb / 3 for nb < 9 (a 16-bit division would be a library call here)
        const int gx = bx + nb - 3 * q - 1, gy = by + q - 1;
        if (gx < 0 || gx >= P.blocks_x || gy < 0 || gy >= P.blocks_y) continue;
        const size_t g = (size_t)gy * P.blocks_x + gx;
        if (lane < 3 * F) {
            r.w[j] = __ldg(P.weights + g * (3 * F) + lane);
            if (f >= NNS) r.s[j] = __ldg(P.mins_inv + g * (2 * NSC) + 2 * (f - NNS) + 1);
        }
        if (lane < F && lane >= NNS) r.m[j] = __ldg(P.mins_inv + g * (2 * NSC) + 2 * (lane - NNS));
    }
    return r;
}
template <int FS>
__device__ __forceinline__ void store_coefficients(const CoefRegs& r, float (*coef)[PT_COEF_S], int warp, int lane, int nwarps) {
    constexpr int F = FeatureSet<FS>::F;
    const int f = lane / 3, c = lane - 3 * f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int nb = warp + j * nwarps;
        if (nb >= 9) continue;
        if (lane < 3 * F) coef[nb][4 * f + c] = __fmul_rn(r.w[j], r.s[j]);  // (a block outside the grid: zeros, never read)
        if (lane < F) coef[nb][4 * lane + 3] = r.m[j];
    }
}
template <int FS>
__device__ __forceinline__ void load_coefficients_scaled(const KParams& P, float (*coef)[PT_COEF_S], int bx, int by, int warp, int lane, int nwarps) {
    store_coefficients<FS>(fetch_coefficients<FS>(P, bx, by, warp, lane, nwarps), coef, warp, lane, nwarps);
}
// the F - 1 non-constant features of a pixel (bmfr.cl:724-741: clean, no noise, no NaN scrub) with the minimum taken off the
// scaled ones (their 1/range lives in the coefficients)
template <int FS>
__device__ __forceinline__ void features_offset(f3 n, f3 p, const float4 (&w)[FeatureSet<FS>::F], float (&feat)[FeatureSet<FS>::F - 1]) {
    constexpr int F = FeatureSet<FS>::F, NSC = FeatureSet<FS>::NSC, NNS = F - NSC;
    int at = 0;
    if (FeatureSet<FS>::NORMALS) { feat[0] = n.x; feat[1] = n.y; feat[2] = n.z; at = 3; }
    const float lin[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int k = 0; k < NSC; ++k) {
        const float mn = w[NNS + k].w;
        feat[at + k] = k < 3 ? __fsub_rn(lin[k], mn) : fmaf(lin[k - 3], lin[k - 3], -mn);
    }
}
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ f3 clamp_negative_nan(f3 c) {  // bmfr.cl:750 (a NaN stays a NaN) in one instruction per component
    return make_f3(max_nan(c.x, 0.f), max_nan(c.y, 0.f), max_nan(c.z, 0.f));
}
// weighted_sum (bmfr.cl:725-750) of two pixels from one set of coefficient loads / of one pixel: the same chain per pixel
template <int FS>
__device__ __forceinline__ void weighted_sum_scaled_px2(f3 n0, f3 p0, f3 n1, f3 p1, const float* __restrict__ cf, f3& out0, f3& out1) {
    constexpr int F = FeatureSet<FS>::F;
    const float4* c4 = reinterpret_cast<const float4*>(cf);
    float4 w[F];
#pragma unroll
    for (int f = 0; f < F; ++f) w[f] = c4[f];
    float f0[F - 1], f1[F - 1];
    features_offset<FS>(n0, p0, w, f0);
    features_offset<FS>(n1, p1, w, f1);
    f3 a = make_f3(w[0].x, w[0].y, w[0].z), b = a;
#pragma unroll
    for (int f = 1; f < F; ++f) {
        a.x = fmaf(w[f].x, f0[f - 1], a.x); a.y = fmaf(w[f].y, f0[f - 1], a.y); a.z = fmaf(w[f].z, f0[f - 1], a.z);
        b.x = fmaf(w[f].x, f1[f - 1], b.x); b.y = fmaf(w[f].y, f1[f - 1], b.y); b.z = fmaf(w[f].z, f1[f - 1], b.z);
    }
    out0 = clamp_negative_nan(a);
    out1 = clamp_negative_nan(b);
}
template <int FS>
__device__ __forceinline__ f3 weighted_sum_scaled_px(f3 n, f3 p, const float* __restrict__ cf) {
    constexpr int F = FeatureSet<FS>::F;
    const float4* c4 = reinterpret_cast<const float4*>(cf);
    float4 w[F];
#pragma unroll
    for (int f = 0; f < F; ++f) w[f] = c4[f];
    float ft[F - 1];
    features_offset<FS>(n, p, w, ft);
    f3 a = make_f3(w[0].x, w[0].y, w[0].z);
#pragma unroll
    for (int f = 1; f < F; ++f) { a.x = fmaf(w[f].x, ft[f - 1], a.x); a.y = fmaf(w[f].y, ft[f - 1], a.y); a.z = fmaf(w[f].z, ft[f - 1], a.z); }
    return clamp_negative_nan(a);
}

struct TapGeom {  // the 2x2 bilinear footprint of one pixel in the previous frame
    int pix, piy;
    float w[4];
};
__device__ __forceinline__ TapGeom tap_geom(float2 pp) {
    TapGeom t;
    t.pix = __float2int_rd(pp.x);
    t.piy = __float2int_rd(pp.y);
    const float frx = __fsub_rn(pp.x, (float)t.pix), fry = __fsub_rn(pp.y, (float)t.piy);
    const float omx = __fsub_rn(1.f, frx), omy = __fsub_rn(1.f, fry);
    t.w[0] = __fmul_rn(omx, omy); t.w[1] = __fmul_rn(frx, omy); t.w[2] = __fmul_rn(omx, fry); t.w[3] = __fmul_rn(frx, fry);
    return t;
}
// accumulate_filtered_data of one pixel (bmfr.cl:778-849) from its four taps (upper row a0[0..5] = two pixels, lower row
// a1), without branches: an unaccepted tap is predicated off as in the reference (its value may be anything), the two
// nested conditions of bmfr.cl:789,834 become selects.  accept = 0 (and frame 0) gives the filtered colour itself.
__device__ __forceinline__ f3 accumulate_taps(const KParams& P, unsigned int accept, unsigned int spp, const float (&w)[4], f3 filtered,
                                              const float* __restrict__ a0, const float* __restrict__ a1) {
    f3 prev = make_f3(0.f, 0.f, 0.f);
    float total = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (accept & (1u << i)) {  // taps are trusted, not re-checked: bmfr.cl:801-832
            const float* t = ((i >> 1) ? a1 : a0) + 3 * (i & 1);
            total = __fadd_rn(total, w[i]);
            prev.x = fmaf(w[i], t[0], prev.x);
            prev.y = fmaf(w[i], t[1], prev.y);
            prev.z = fmaf(w[i], t[2], prev.z);
        }
    }
    const bool any = total > 0.f;
    const float alpha = any ? fmaxf(fast_rcp((float)spp), P.second_blend_alpha) : 1.f;  // bmfr.cl:838-839
    const float k = any ? __fmul_rn(__fsub_rn(1.f, alpha), fast_rcp(total)) : 0.f;       // (1 - alpha) / total_weight
    return make_f3(fmaf(alpha, filtered.x, __fmul_rn(k, prev.x)), fmaf(alpha, filtered.y, __fmul_rn(k, prev.y)),
                   fmaf(alpha, filtered.z, __fmul_rn(k, prev.z)));
}
// clamp(powr(max(0, a * b), 0.454545f), 0, 1), bmfr.cl:852-856, with the clamp in front of the power (x -> x^0.4545 is
// monotonic and maps [0, 1] onto itself): the saturating multiply is one instruction, and a NaN product becomes 0 like
// fmax(0.f, NaN) in OpenCL.  0 -> lg2 = -inf -> ex2 = 0, like powr(0, y > 0).
__device__ __forceinline__ float tone_map_product(float a, float b) {
    float v, l, e;
    asm("mul.rn.ftz.sat.f32 %0, %1, %2;" : "=f"(v) : "f"(a), "f"(b));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(v));
    l = __fmul_rn(l, 0.454545f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(l));
    return e;
}
__device__ __forceinline__ f3 to_ycocg_rn(f3 c) {  // bmfr.cl:184-190: (r + 2g + b, 2r - 2b, -r + 2g - b) in five operations
    const float t = __fadd_rn(c.x, c.z), d = __fsub_rn(c.x, c.z);
    return make_f3(fmaf(2.f, c.y, t), __fadd_rn(d, d), fmaf(2.f, c.y, -t));
}
// bmfr.cl:192-198 for a colour that already carries the factor 0.25 (the blend of phase B folds it into its weights)
__device__ __forceinline__ f3 from_quarter_ycocg(f3 c) {
    const float u = __fsub_rn(c.x, c.z);
    return make_f3(__fadd_rn(u, c.y), __fadd_rn(c.x, c.z), __fsub_rn(u, c.y));
}
__device__ __forceinline__ f3 tone_ycocg(f3 alb, f3 accum) {
    return to_ycocg_rn(make_f3(tone_map_product(alb.x, accum.x), tone_map_product(alb.y, accum.y), tone_map_product(alb.z, accum.z)));
}
// Bilinear sample of the previous TAA result in YCoCg (bmfr.cl:922-965) from taps in registers; bit i of ok: tap i lies in
// the image (ALL: every one does).  With all four taps the weights sum to 1 +- 2 ulp and the division of bmfr.cl:962 is
// left out; with fewer, 0 * inf = NaN on the image edge like its 0 / 0.
template <bool ALL>
__device__ __forceinline__ f3 history_ycc(const float (&w)[4], const float* __restrict__ r0, const float* __restrict__ r1, unsigned int ok) {
    f3 hp = make_f3(0.f, 0.f, 0.f);
    float total = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bmfr.cl:929-960
        if (ALL || (ok & (1u << i))) {
            const float* t = ((i >> 1) ? r1 : r0) + 3 * (i & 1);
            hp.x = fmaf(w[i], t[0], hp.x);
            hp.y = fmaf(w[i], t[1], hp.y);
            hp.z = fmaf(w[i], t[2], hp.z);
            if (!ALL) total = __fadd_rn(total, w[i]);
        }
    }
    if (!ALL && ok != 15u) {
        const float inv = fast_rcp(total);
        hp = make_f3(__fmul_rn(hp.x, inv), __fmul_rn(hp.y, inv), __fmul_rn(hp.z, inv));
    }
    return to_ycocg_rn(hp);
}
// which taps of the history sample lie in the image (bmfr.cl:929-960), and does the pixel take the temporal path (bmfr.cl:884-890)?
__device__ __forceinline__ unsigned int history_taps_ok(const KParams& P, const TapGeom& t) {
    const bool x0 = t.pix >= 0, x1 = t.pix < P.W - 1, y0 = t.piy >= 0, y1 = t.piy < P.H - 1;
    return (x0 && y0 ? 1u : 0u) | (x1 && y0 ? 2u : 0u) | (x0 && y1 ? 4u : 0u) | (x1 && y1 ? 8u : 0u);
}
__device__ __forceinline__ bool history_in_reach(const KParams& P, const TapGeom& t) {
    return !(t.pix < -1 || t.piy < -1 || t.pix >= P.W || t.piy >= P.H);
}
__device__ __forceinline__ void f3_to(float* d, f3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

// The accumulated filtered colour and the TAA result ("history": gathered from by the next frame, four taps per pixel and
// buffer) are HS floats per pixel.  HS = 3 is the reference's layout (bmfr.cl:224-241) and the default.  HS = 4 (tuning switch
// BMFR_HISTORY_PADDED in bmfr_pipeline.cu: whole-image FUSED contexts whose post pass runs this kernel, KParams::hist_stride)
// pads a pixel to 16 bytes: a tap is ONE 128-bit load whose warp-wide footprint is four cache lines in four L1 wavefronts,
// where three 32-bit loads at a 12-byte stride take ten to twelve.  Measured at 1080p with every GPU test green: 60.2 against
// 55.8 us — the 16 B per pixel of extra DRAM traffic (two reads, two writes of the padding, +17 %) cost more than the 62 L1
// wavefronts per pixel it saves: the kernel follows its bytes, not its L1 requests.  Kept, off.
template <int HS>
__device__ __forceinline__ f3 load_hist(const float* __restrict__ b, unsigned int i) {
    if (HS == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(b) + i);
        return make_f3(v.x, v.y, v.z);
    }
    return load_f3(b, i);
}
template <int HS>
__device__ __forceinline__ void store_hist(float* __restrict__ b, unsigned int i, f3 v) {
    if (HS == 4) reinterpret_cast<float4*>(b)[i] = make_float4(v.x, v.y, v.z, 0.f);
    else store_f3(b, i, v);
}
// six floats = the pixels i and i + 1 of one tap row
template <int HS>
__device__ __forceinline__ void load_hist_pair(float* __restrict__ d, const float* __restrict__ row, unsigned int i) {
    if (HS == 4) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(row) + i), v = __ldg(reinterpret_cast<const float4*>(row) + i + 1);
        d[0] = u.x; d[1] = u.y; d[2] = u.z; d[3] = v.x; d[4] = v.y; d[5] = v.z;
    } else {
        const float* p = row + (size_t)(i * 3u);
#pragma unroll
        for (int k = 0; k < 6; ++k) d[k] = __ldg(p + k);
    }
}

// One pixel on its own — a ring pixel, or a pixel of a pair that a strip or image edge cuts: taps fetched from clamped
// addresses (an accepted tap is in the image, bmfr.cl:386-392, so its clamped address is its own), then the shared arithmetic.
// store: the pixel belongs to the tile (accumulated colour written, history sampled if owned); returns the temporal flag.
template <bool STRIP, int FS, int HS>
__device__ __forceinline__ bool single_pixel(PostStage& sh, const KParams& P, const TileGeom& G, const AlbGate& gate, const float* cf, int hx, int hy,
                                             int x, int y, bool store, bool own, f3& hist, bool zone) {
    const PixelIn in = load_pixel_staged(sh, P, G, hx, hy, x, y);
    const f3 filtered = weighted_sum_scaled_px<FS>(in.n, in.p, cf);
    const TapGeom t = tap_geom(in.pp);
    const bool want_hist = store && own && P.frame > 0 && history_in_reach(P, t);
    float a0[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a1[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, r0[6], r1[6];
    unsigned int accept = 0;
    hist = make_f3(0.f, 0.f, 0.f);
    if (P.frame > 0) {
        accept = in.accept;
        const int rlo = STRIP ? P.state2_row0 : 0, rhi = (STRIP ? P.state2_row1 : P.H) - 1;
        const int cx[2] = {min(max(t.pix, 0), P.W - 1), min(max(t.pix + 1, 0), P.W - 1)};
        const int cy[2] = {min(max(t.piy, rlo), rhi), min(max(t.piy + 1, rlo), rhi)};
        const unsigned int ok = history_taps_ok(P, t);
        if (STRIP) {  // a wanted tap whose row this strip does not hold: reported once, the run is invalid
            const bool out0 = t.piy < P.state2_row0 || t.piy >= P.state2_row1, out1 = t.piy + 1 < P.state2_row0 || t.piy + 1 >= P.state2_row1;
            const unsigned int wanted = accept | (want_hist ? ok : 0u);
            if (((wanted & 3u) != 0 && out0) || ((wanted & 12u) != 0 && out1)) *P.oob_flag = 1;
        }
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const unsigned int l0 = pix_index(P, cx[dx], cy[0]), l1 = pix_index(P, cx[dx], cy[1]);
            f3_to(a0 + 3 * dx, load_hist<HS>(P.accum_prev, l0));
            f3_to(a1 + 3 * dx, load_hist<HS>(P.accum_prev, l1));
            if (want_hist) {
                f3_to(r0 + 3 * dx, load_hist<HS>(P.result_prev, l0));
                f3_to(r1 + 3 * dx, load_hist<HS>(P.result_prev, l1));
            }
        }
        if (want_hist) hist = history_ycc<false>(t.w, r0, r1, ok);
    }
    const f3 accum = accumulate_taps(P, accept, in.spp, t.w, filtered, a0, a1);
    if (store) {
        store_hist<HS>(P.accum_cur, in.lp, accum);
        if (STRIP && zone) stage_accum(sh, G, hx, hy, accum);
    }
    put_ycc_i(sh, P, G, hx, hy, x, y, tone_ycocg(albedo_cell(sh, G, gate, hx, hy), accum));
    return want_hist;
}

#ifndef BMFR_POST_FAST
#define BMFR_POST_FAST 1
#endif
#ifndef BMFR_POST_TMA_MIN_BLOCKS
#define BMFR_POST_TMA_MIN_BLOCKS 3  // CTAs per SM with four pixels per thread (256 threads); two pixels per thread: 512 threads, two CTAs
#endif
#ifndef BMFR_POST_PX
#define BMFR_POST_PX 4  // pixels per thread of the whole-image instantiation (4: 8 warps per tile, 2: 16)
#endif
#ifndef BMFR_POST_HISTORY_PREFETCH
#define BMFR_POST_HISTORY_PREFETCH 0  // margin in pixels around the tile (0: off)
#endif
#ifndef BMFR_POST_RING_SPREAD
#define BMFR_POST_RING_SPREAD 0
#endif
#ifndef BMFR_POST_SMEM_PAD
#define BMFR_POST_SMEM_PAD 0  // extra dynamic shared memory per CTA of the whole-image instantiation: fewer CTAs per SM, more L1
#endif

// One CTA per 32 x 32 tile; thread (lane, warp) owns the column strip x = x0 + lane, rows PX * warp .. + PX - 1.
// Phase A of one tile: filtered -> accumulated -> tone-mapped colour of the tile's pixels and of its ring (YCoCg values into the
// albedo cells), the TAA history samples of the tile's pixels (hist) and their flags (live: bit s = pixel s is written, bit
// 4 + s = it takes the temporal path).  The caller puts a CTA barrier behind it.
template <bool STRIP, int FS, int PX, int HS>
__device__ __forceinline__ void post_phase_a(PostStage& sh, const KParams& P, const TileGeom& G, const AlbGate& gate, const float (*coef)[PT_COEF_S],
                                             int tid, int lane, int warp, bool zone, f3 (&hist)[PX], unsigned int& live) {
    const int x = G.x0 + lane;
    const int rlo = STRIP ? P.state2_row0 : 0, rhi = (STRIP ? P.state2_row1 : P.H) - 1;  // rows of accum / result this context holds
    // interior tile (CTA-uniform): the halo lies inside the image and inside the rows phase A covers, every pixel is owned,
    // no strip edge is near (no zone duties), frame > 0
    const bool interior = BMFR_POST_FAST != 0 && P.frame > 0 && !zone && G.x0 >= 1 && G.x0 + 33 <= P.W && G.y0 >= 1 && G.y0 + 33 <= P.H &&
                          G.y0 - 1 >= P.py0 && G.y0 + 33 <= P.py1 && G.y0 >= P.own_y0 && G.y0 + 32 <= P.own_y1;
    live = 0;
    if (interior) {
        const int WH = HS * P.W;  // floats per row of the history buffers
#pragma unroll
        for (int s = 0; s < PX; s += 2) {
            const int ty = PX * warp + s, y = G.y0 + ty;
            const PixelIn i0 = load_pixel_staged(sh, P, G, lane + 1, ty + 1, x, y), i1 = load_pixel_staged(sh, P, G, lane + 1, ty + 2, x, y + 1);
            const TapGeom g0 = tap_geom(i0.pp), g1 = tap_geom(i1.pp);
            // the pair's footprints: stacked (rows piy, piy + 1, piy + 2 of columns pix, pix + 1) and inside the image / the rows held?
            const bool easy = (unsigned int)g0.pix < (unsigned int)(P.W - 1) && g0.piy >= rlo && g0.piy + 2 <= rhi && g1.pix == g0.pix &&
                              g1.piy == g0.piy + 1;
            float A[3][6], R[3][6];
            f3 fl0, fl1;
            if (__all_sync(0xffffffffu, easy)) {
                const unsigned int base = pix_index(P, g0.pix, g0.piy);
#pragma unroll
                for (int row = 0; row < 3; ++row) {
                    load_hist_pair<HS>(A[row], P.accum_prev + (size_t)row * WH, base);
                    load_hist_pair<HS>(R[row], P.result_prev + (size_t)row * WH, base);
                }
                weighted_sum_scaled_px2<FS>(i0.n, i0.p, i1.n, i1.p, coef[4], fl0, fl1);  // its arithmetic overlaps the gathers
                hist[s] = history_ycc<true>(g0.w, R[0], R[1], 15u);
                hist[s + 1] = history_ycc<true>(g1.w, R[1], R[2], 15u);
                live |= 0x33u << s;
            } else {  // some lane's footprint touches an edge or its pair is not stacked: clamped addresses, per-tap validity
                const int cx0[2] = {min(max(g0.pix, 0), P.W - 1), min(max(g0.pix + 1, 0), P.W - 1)};
                const int cx1[2] = {min(max(g1.pix, 0), P.W - 1), min(max(g1.pix + 1, 0), P.W - 1)};
                const int ry0[2] = {min(max(g0.piy, rlo), rhi), min(max(g0.piy + 1, rlo), rhi)};
                const int ry1[2] = {min(max(g1.piy, rlo), rhi), min(max(g1.piy + 1, rlo), rhi)};
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const unsigned int l0 = pix_index(P, cx0[dx], ry0[0]), l1 = pix_index(P, cx0[dx], ry0[1]), l2 = pix_index(P, cx1[dx], ry1[1]);
                    f3_to(&A[0][3 * dx], load_hist<HS>(P.accum_prev, l0)); f3_to(&A[1][3 * dx], load_hist<HS>(P.accum_prev, l1)); f3_to(&A[2][3 * dx], load_hist<HS>(P.accum_prev, l2));
                    f3_to(&R[0][3 * dx], load_hist<HS>(P.result_prev, l0)); f3_to(&R[1][3 * dx], load_hist<HS>(P.result_prev, l1)); f3_to(&R[2][3 * dx], load_hist<HS>(P.result_prev, l2));
                }
                weighted_sum_scaled_px2<FS>(i0.n, i0.p, i1.n, i1.p, coef[4], fl0, fl1);
                const bool t0 = history_in_reach(P, g0), t1 = history_in_reach(P, g1);
                const unsigned int ok0 = history_taps_ok(P, g0), ok1 = history_taps_ok(P, g1);
                if (STRIP) {
                    const bool o00 = g0.piy < P.state2_row0 || g0.piy >= P.state2_row1, o01 = g0.piy + 1 < P.state2_row0 || g0.piy + 1 >= P.state2_row1;
                    const bool o10 = g1.piy < P.state2_row0 || g1.piy >= P.state2_row1, o11 = g1.piy + 1 < P.state2_row0 || g1.piy + 1 >= P.state2_row1;
                    const unsigned int w0 = i0.accept | (t0 ? ok0 : 0u), w1 = i1.accept | (t1 ? ok1 : 0u);
                    if (((w0 & 3u) != 0 && o00) || ((w0 & 12u) != 0 && o01) || ((w1 & 3u) != 0 && o10) || ((w1 & 12u) != 0 && o11)) *P.oob_flag = 1;
                }
                hist[s] = history_ycc<false>(g0.w, R[0], R[1], ok0);
                if (!(cx1[0] == cx0[0] && cx1[1] == cx0[1] && ry1[0] == ry0[1])) {  // footprints not stacked: fetch the lower pixel's upper row
                    // (the upper pixel's accumulation below needs its own lower row: finish it first)
                    const f3 acc0 = accumulate_taps(P, i0.accept, i0.spp, g0.w, fl0, A[0], A[1]);
                    store_hist<HS>(P.accum_cur, i0.lp, acc0);
                    put_cell_i(sh, G, lane + 1, ty + 1, tone_ycocg(albedo_cell(sh, G, gate, lane + 1, ty + 1), acc0));
#pragma unroll
                    for (int dx = 0; dx < 2; ++dx) {
                        const unsigned int l = pix_index(P, cx1[dx], ry1[0]);
                        f3_to(&A[1][3 * dx], load_hist<HS>(P.accum_prev, l));
                        f3_to(&R[1][3 * dx], load_hist<HS>(P.result_prev, l));
                    }
                    hist[s + 1] = history_ycc<false>(g1.w, R[1], R[2], ok1);
                    const f3 acc1 = accumulate_taps(P, i1.accept, i1.spp, g1.w, fl1, A[1], A[2]);
                    store_hist<HS>(P.accum_cur, i1.lp, acc1);
                    put_cell_i(sh, G, lane + 1, ty + 2, tone_ycocg(albedo_cell(sh, G, gate, lane + 1, ty + 2), acc1));
                    live |= ((1u | (t0 ? 16u : 0u)) | ((1u | (t1 ? 16u : 0u)) << 1)) << s;
                    continue;
                }
                hist[s + 1] = history_ycc<false>(g1.w, R[1], R[2], ok1);
                live |= ((1u | (t0 ? 16u : 0u)) | ((1u | (t1 ? 16u : 0u)) << 1)) << s;
            }
            const f3 acc0 = accumulate_taps(P, i0.accept, i0.spp, g0.w, fl0, A[0], A[1]);
            const f3 acc1 = accumulate_taps(P, i1.accept, i1.spp, g1.w, fl1, A[1], A[2]);
            store_hist<HS>(P.accum_cur, i0.lp, acc0);
            store_hist<HS>(P.accum_cur, i1.lp, acc1);
            put_cell_i(sh, G, lane + 1, ty + 1, tone_ycocg(albedo_cell(sh, G, gate, lane + 1, ty + 1), acc0));
            put_cell_i(sh, G, lane + 1, ty + 2, tone_ycocg(albedo_cell(sh, G, gate, lane + 1, ty + 2), acc1));
        }
    } else {
        // general tile: image edges, strip edges, rows that are only partly covered or owned, zone duties, frame 0
        const bool col_ok = x >= 0 && x < P.W;
#pragma unroll
        for (int s = 0; s < PX; s += 2) {
            const int ty = PX * warp + s, y = G.y0 + ty;
            hist[s] = hist[s + 1] = make_f3(0.f, 0.f, 0.f);
            const bool v0 = col_ok && y >= P.py0 && y < P.py1, v1 = col_ok && y + 1 >= P.py0 && y + 1 < P.py1;
            const bool own0 = y >= P.own_y0 && y < P.own_y1, own1 = y + 1 >= P.own_y0 && y + 1 < P.own_y1;
            if (v0 && v1) {
                const PixelIn i0 = load_pixel_staged(sh, P, G, lane + 1, ty + 1, x, y), i1 = load_pixel_staged(sh, P, G, lane + 1, ty + 2, x, y + 1);
                const TapGeom g0 = tap_geom(i0.pp), g1 = tap_geom(i1.pp);
                float A[3][6], R[3][6];
                f3 fl0, fl1;
                bool t0 = false, t1 = false;
                unsigned int acc_bits0 = 0, acc_bits1 = 0;
                bool stacked = true;
                if (P.frame > 0) {
                    // one round of independent loads from clamped addresses: tap rows R0, R1 of the upper pixel, R2 of the lower;
                    // the lower pixel's upper row is R1 when the two footprints are stacked (checked below)
                    const int cx0[2] = {min(max(g0.pix, 0), P.W - 1), min(max(g0.pix + 1, 0), P.W - 1)};
                    const int cx1[2] = {min(max(g1.pix, 0), P.W - 1), min(max(g1.pix + 1, 0), P.W - 1)};
                    const int ry0[2] = {min(max(g0.piy, rlo), rhi), min(max(g0.piy + 1, rlo), rhi)};
                    const int ry1[2] = {min(max(g1.piy, rlo), rhi), min(max(g1.piy + 1, rlo), rhi)};
#pragma unroll
                    for (int dx = 0; dx < 2; ++dx) {
                        const unsigned int l0 = pix_index(P, cx0[dx], ry0[0]), l1 = pix_index(P, cx0[dx], ry0[1]), l2 = pix_index(P, cx1[dx], ry1[1]);
                        f3_to(&A[0][3 * dx], load_hist<HS>(P.accum_prev, l0)); f3_to(&A[1][3 * dx], load_hist<HS>(P.accum_prev, l1)); f3_to(&A[2][3 * dx], load_hist<HS>(P.accum_prev, l2));
                        f3_to(&R[0][3 * dx], load_hist<HS>(P.result_prev, l0)); f3_to(&R[1][3 * dx], load_hist<HS>(P.result_prev, l1)); f3_to(&R[2][3 * dx], load_hist<HS>(P.result_prev, l2));
                    }
                    weighted_sum_scaled_px2<FS>(i0.n, i0.p, i1.n, i1.p, coef[4], fl0, fl1);  // its arithmetic overlaps the gathers
                    acc_bits0 = i0.accept; acc_bits1 = i1.accept;
                    t0 = own0 && history_in_reach(P, g0);
                    t1 = own1 && history_in_reach(P, g1);
                    const unsigned int ok0 = history_taps_ok(P, g0), ok1 = history_taps_ok(P, g1);
                    if (STRIP) {  // a wanted tap whose row this strip does not hold: reported once, the run is invalid
                        const bool o00 = g0.piy < P.state2_row0 || g0.piy >= P.state2_row1, o01 = g0.piy + 1 < P.state2_row0 || g0.piy + 1 >= P.state2_row1;
                        const bool o10 = g1.piy < P.state2_row0 || g1.piy >= P.state2_row1, o11 = g1.piy + 1 < P.state2_row0 || g1.piy + 1 >= P.state2_row1;
                        const unsigned int w0 = i0.accept | (t0 ? ok0 : 0u), w1 = i1.accept | (t1 ? ok1 : 0u);
                        if (((w0 & 3u) != 0 && o00) || ((w0 & 12u) != 0 && o01) || ((w1 & 3u) != 0 && o10) || ((w1 & 12u) != 0 && o11)) *P.oob_flag = 1;
                    }
                    if (t0) hist[s] = history_ycc<false>(g0.w, R[0], R[1], ok0);
                    stacked = cx1[0] == cx0[0] && cx1[1] == cx0[1] && ry1[0] == ry0[1];
                    if (!stacked) {  // rare: finish the upper pixel, then fetch the lower pixel's own upper row
                        const f3 acc0 = accumulate_taps(P, acc_bits0, i0.spp, g0.w, fl0, A[0], A[1]);
                        store_hist<HS>(P.accum_cur, i0.lp, acc0);
                        if (STRIP && zone) stage_accum(sh, G, lane + 1, ty + 1, acc0);
                        put_ycc_i(sh, P, G, lane + 1, ty + 1, x, y, tone_ycocg(albedo_cell(sh, G, gate, lane + 1, ty + 1), acc0));
#pragma unroll
                        for (int dx = 0; dx < 2; ++dx) {
                            const unsigned int l = pix_index(P, cx1[dx], ry1[0]);
                            f3_to(&A[1][3 * dx], load_hist<HS>(P.accum_prev, l));
                            f3_to(&R[1][3 * dx], load_hist<HS>(P.result_prev, l));
                        }
                    }
                    if (t1) hist[s + 1] = history_ycc<false>(g1.w, R[1], R[2], ok1);
                } else {  // frame 0: no temporal path (bmfr.cl:784, 884): accept = 0 makes the accumulated colour the filtered one
#pragma unroll
                    for (int row = 0; row < 3; ++row)
#pragma unroll
                        for (int k = 0; k < 6; ++k) A[row][k] = 0.f;
                    weighted_sum_scaled_px2<FS>(i0.n, i0.p, i1.n, i1.p, coef[4], fl0, fl1);
                }
                if (stacked) {
                    const f3 acc0 = accumulate_taps(P, acc_bits0, i0.spp, g0.w, fl0, A[0], A[1]);
                    store_hist<HS>(P.accum_cur, i0.lp, acc0);
                    if (STRIP && zone) stage_accum(sh, G, lane + 1, ty + 1, acc0);
                    put_ycc_i(sh, P, G, lane + 1, ty + 1, x, y, tone_ycocg(albedo_cell(sh, G, gate, lane + 1, ty + 1), acc0));
                }
                const f3 acc1 = accumulate_taps(P, acc_bits1, i1.spp, g1.w, fl1, A[1], A[2]);
                store_hist<HS>(P.accum_cur, i1.lp, acc1);
                if (STRIP && zone) stage_accum(sh, G, lane + 1, ty + 2, acc1);
                put_ycc_i(sh, P, G, lane + 1, ty + 2, x, y + 1, tone_ycocg(albedo_cell(sh, G, gate, lane + 1, ty + 2), acc1));
                live |= ((own0 ? 1u : 0u) | (t0 ? 16u : 0u)) << s;
                live |= ((own1 ? 1u : 0u) | (t1 ? 16u : 0u)) << (s + 1);
            } else if (v0) {  // a strip or image edge cuts the pair
                const bool t = single_pixel<STRIP, FS, HS>(sh, P, G, gate, coef[4], lane + 1, ty + 1, x, y, true, own0, hist[s], zone);
                live |= ((own0 ? 1u : 0u) | (t ? 16u : 0u)) << s;
            } else if (v1) {
                const bool t = single_pixel<STRIP, FS, HS>(sh, P, G, gate, coef[4], lane + 1, ty + 2, x, y + 1, true, own1, hist[s + 1], zone);
                live |= ((own1 ? 1u : 0u) | (t ? 16u : 0u)) << (s + 1);
            }
        }
    }
    // phase A, ring: 4 * 33 = 132 pixels, coefficients of the pixel's own block.  BMFR_POST_RING_SPREAD: 17 per warp (with
    // four pixels per thread) instead of the first 132 threads — every warp then reaches the phase barrier after the same
    // work, at the price of the ring code being issued by all warps
    const int rid = BMFR_POST_RING_SPREAD && PX == 4 ? (lane < 17 ? warp * 17 + lane : 4 * (PT_HALO - 1)) : tid;
    if (rid < 4 * (PT_HALO - 1)) {
        int hx, hy;
        if (rid < PT_HALO - 1) { hx = rid; hy = 0; }
        else if (rid < 2 * (PT_HALO - 1)) { hx = PT_HALO - 1; hy = rid - (PT_HALO - 1); }
        else if (rid < 3 * (PT_HALO - 1)) { hx = PT_HALO - 1 - (rid - 2 * (PT_HALO - 1)); hy = PT_HALO - 1; }
        else { hx = 0; hy = PT_HALO - 1 - (rid - 3 * (PT_HALO - 1)); }
        const int rx = G.x0 + hx - 1, ry = G.y0 + hy - 1;
        if (interior || (rx >= 0 && rx < P.W && ry >= P.py0 && ry < P.py1)) {
            const int nb = ((hy == 0) ? 0 : (hy == PT_HALO - 1) ? 6 : 3) + ((hx == 0) ? 0 : (hx == PT_HALO - 1) ? 2 : 1);
            f3 unused;
            single_pixel<STRIP, FS, HS>(sh, P, G, gate, coef[nb], hx, hy, rx, ry, false, false, unused, zone);
        }
    }
}

// Phase B of one tile (after the barrier): neighbourhood clamp of the history samples, blend, TAA result stored.
template <bool STRIP, int PX, int HS>
__device__ __forceinline__ void post_phase_b(PostStage& sh, const KParams& P, const TileGeom& G, int lane, int warp, bool zone, const f3 (&hist)[PX],
                                             unsigned int live) {
    const int x = G.x0 + lane;
    // phase B: clamp the history samples to the neighbourhood box, component by component (bmfr.cl:893-920, 967-969), blend
    // (bmfr.cl:971-973; in YCoCg, the weights carry the 0.25 of the conversion back) and store.  Halo rows PX * warp ..
    // PX * warp + PX + 1 cover the 3x3 neighbourhoods of the strip; this thread's column is lane + 1.  A pixel on the
    // copy-through path of bmfr.cl:884-890 keeps its own colour.
    const float a = __fmul_rn(0.25f, P.taa_blend_alpha), oma = __fmul_rn(0.25f, __fsub_rn(1.f, P.taa_blend_alpha));
    f3 out[PX];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float ctr[PX + 2], rmin[PX + 2], rmax[PX + 2];
#pragma unroll
        for (int r = 0; r < PX + 2; ++r) {
            const float* row = &sh.alb[PX * warp + r][G.sh_rgb + c];
            const float l = row[3 * lane], m = row[3 * lane + 3], rr = row[3 * lane + 6];
            ctr[r] = m;
            rmin[r] = fminf(fminf(l, m), rr);
            rmax[r] = fmaxf(fmaxf(l, m), rr);
        }
#pragma unroll
        for (int s = 0; s < PX; ++s) {
            const float min_box = fminf(fminf(rmin[s], rmin[s + 1]), rmin[s + 2]);
            const float max_box = fmaxf(fmaxf(rmax[s], rmax[s + 1]), rmax[s + 2]);
            const float min_cross = fminf(fminf(ctr[s], rmin[s + 1]), ctr[s + 2]);
            const float max_cross = fmaxf(fmaxf(ctr[s], rmax[s + 1]), ctr[s + 2]);
            const float lo = __fmul_rn(__fadd_rn(min_box, min_cross), 0.5f), hi = __fmul_rn(__fadd_rn(max_box, max_cross), 0.5f);
            const float h = (c == 0) ? hist[s].x : (c == 1) ? hist[s].y : hist[s].z;
            const float hc = fminf(fmaxf(h, lo), hi);
            const float mine = ctr[s + 1];
            const float v = (live & (16u << s)) ? fmaf(a, mine, __fmul_rn(oma, hc)) : __fmul_rn(0.25f, mine);
            if (c == 0) out[s].x = v; else if (c == 1) out[s].y = v; else out[s].z = v;
        }
    }
#pragma unroll
    for (int s = 0; s < PX; ++s) {
        if (!(live & (1u << s))) continue;
        const unsigned int lp = pix_index(P, x, G.y0 + PX * warp + s);
        const f3 rgb = from_quarter_ycocg(out[s]);
        store_hist<HS>(P.result_cur, lp, rgb);
        if (P.user_out) store_f3(P.user_out, lp, rgb);
        if (STRIP && zone) stage_result(sh, G, lane + 1, PX * warp + s + 1, rgb);
    }
}

template <bool STRIP, int FS, int PX, int HS>
__global__ void __launch_bounds__(1024 / PX, PX == 4 ? BMFR_POST_TMA_MIN_BLOCKS : 2) post_tma_kernel(const __grid_constant__ KParams P, const __grid_constant__ PostMaps M) {
    static_assert(PX == 4 || PX == 2, "pixels per thread: two pairs or one");
    constexpr int THREADS = 1024 / PX, WARPS = 32 / PX;
    extern __shared__ __align__(128) unsigned char post_smem[];
    PostStage& sh = *reinterpret_cast<PostStage*>(post_smem);
    const int bx = blockIdx.x, by = P.by0 + (STRIP ? halo_row_order(P.halo_p, blockIdx.y, gridDim.y) : sweep_row(P, blockIdx.y, gridDim.y));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    TileGeom G;
    G.x0 = bx * 32 - 16 + P.off_x;
    G.y0 = by * 32 - 16 + P.off_y;
    const int f_rgb = 3 * (G.x0 - 1), f_pp = 2 * (G.x0 - 1), b_u8 = G.x0 - 1;
    const int c_rgb = f_rgb & ~3, c_pp = f_pp & ~3, c_u8 = b_u8 & ~15;  // 16-byte aligned box starts (floor, also for negatives)
    G.sh_rgb = f_rgb - c_rgb; G.sh_pp = f_pp - c_pp; G.sh_u8 = b_u8 - c_u8;

    if (tid == 0) {
        mbar_init(&sh.bar, 1);
        mbar_fence_init();
        // Requested before the grid dependency is resolved: the caller's inputs and the reprojection's outputs are
        // complete by now — in the in-order stream this grid's CTAs start only after every CTA of the fit has passed
        // its own wait for the reprojection (which waited for the caller's producer), in the overlapped mode the fit's
        // event already orders this launch.  Rows / columns outside the image (or the strip) arrive as zeros.
        const int c1 = G.y0 - 1 - P.row0;
        mbar_expect_tx(&sh.bar, PostStage::TX);
#if BMFR_L2_HINTS >= 2  // last readers of the reprojection's per-pixel outputs, only reader of the albedo
        tma_load_tile_hint(&sh.pp[0][0], &M.pp, c_pp, c1, &sh.bar, BMFR_L2_ONCE);
        tma_load_tile_hint(&sh.acc[0][0], &M.accept, c_u8, c1, &sh.bar, BMFR_L2_ONCE);
        tma_load_tile(&sh.spp[0][0], &M.spp, c_u8, c1, &sh.bar);
        tma_load_tile(&sh.nrm[0][0], &M.normals, c_rgb, c1, &sh.bar);
        tma_load_tile(&sh.pos[0][0], &M.positions, c_rgb, c1, &sh.bar);
        tma_load_tile_hint(&sh.alb[0][0], &M.albedo, c_rgb, c1, &sh.bar, BMFR_L2_ONCE);
#else
        tma_load_tile(&sh.pp[0][0], &M.pp, c_pp, c1, &sh.bar);
        tma_load_tile(&sh.acc[0][0], &M.accept, c_u8, c1, &sh.bar);
        tma_load_tile(&sh.spp[0][0], &M.spp, c_u8, c1, &sh.bar);
        tma_load_tile(&sh.nrm[0][0], &M.normals, c_rgb, c1, &sh.bar);
        tma_load_tile(&sh.pos[0][0], &M.positions, c_rgb, c1, &sh.bar);
        tma_load_tile(&sh.alb[0][0], &M.albedo, c_rgb, c1, &sh.bar);
#endif
    }
    // strips: a CTA near a strip edge waits for the neighbours' accumulated colour / TAA rows of the previous frame; its
    // first look at the flags is in flight across the wait for the fit
    const bool zone = STRIP && halo_in_zone(P.halo_p, G.y0 - 1, G.y0 + PT_TILE + 1);
    const HaloPeek peek = halo_peek(P.halo_p, zone);
#if BMFR_POST_HISTORY_PREFETCH > 0
    // The two tap gathers of phase A are this kernel's exposed latency (long-scoreboard stalls: 3.8 per issued instruction,
    // 2/3 of their L1 misses also miss the L2: the previous frame's accumulated colour and TAA result were written a frame
    // ago).  Where the taps will land is only known once the tile's previous-pixel positions have arrived, but it is
    // almost always near the tile itself: start the DRAM -> L2 fetch of the tile's own area (+ a margin) of both buffers
    // now, a tile-load latency ahead of the gathers.  Lines a neighbouring tile asks for as well are fetched once.
    if (P.frame > 0 && !(STRIP && zone)) {
        constexpr int MARGIN = BMFR_POST_HISTORY_PREFETCH;
        const int hr0 = STRIP ? P.state2_row0 : 0, hr1 = STRIP ? P.state2_row1 : P.H;
        const int ya = max(G.y0 - MARGIN, hr0), yb = min(G.y0 + PT_TILE + MARGIN, hr1);
        const int ba = (max(G.x0 - MARGIN, 0) * 12) & ~127, bb = min(G.x0 + PT_TILE + MARGIN, P.W) * 12;  // byte range of a row
        const int lines = (bb - ba + 127) >> 7;
        const int total = (yb - ya) * lines * 2;
        for (int i = tid; i < total; i += THREADS) {
            const int j = i >> 1, r = j / lines, l = j - r * lines;
            const char* base = reinterpret_cast<const char*>((i & 1) ? P.result_prev : P.accum_prev);
            prefetch_l2(base + ((size_t)(ya + r - P.row0) * P.W * 12 + ba + l * 128));
        }
    }
#endif
    pdl_wait();     // the fit of this frame is complete (weights, min/max)
    pdl_trigger();  // only now, so that "this frame's fit and reprojection are complete" also holds for the successor
    stamp_begin(P, 2);
    if (zone) halo_poll(P.halo_p, peek);

    load_coefficients_scaled<FS>(P, sh.coef[0], bx, by, warp, lane, WARPS);
    __syncthreads();  // the coefficients and the barrier's initialisation are visible
    mbar_wait_hot(&sh.bar, 0);

    f3 hist[PX];
    unsigned int live;
    const AlbGate gate{nullptr, 0u};  // the albedo came with the other inputs
    post_phase_a<STRIP, FS, PX, HS>(sh, P, G, gate, sh.coef[0], tid, lane, warp, zone, hist, live);
    __syncthreads();
    post_phase_b<STRIP, PX, HS>(sh, P, G, lane, warp, zone, hist, live);
    if (STRIP && zone) {
        __syncthreads();  // the staged rows are complete
        post_push_rows<THREADS>(P, sh, G, tid);
        halo_finish(P.halo_p, halo_cta_pushes(P.halo_p, G.y0, G.y0 + PT_TILE));
    }
    stamp_end(P, 2);
}

// ------------------------------------------------------------------------------------------------
// Persistent variant (BMFR_POST_PERSIST, the default): min(tiles, 3 x SMs) CTAs, each walks over tiles t = blockIdx.x,
// + gridDim.x, ... of the launch.  What it buys: a tile's inputs no longer arrive while its CTA sits idle — the source-
// level profile of the one-tile-per-CTA kernel has 13 % of the warp time in that wait (tile + coefficients) on top of
// the launch of 2135 CTAs.  The stage is not doubled (three CTAs per SM must keep fitting, and the L1 beside them keeps its
// 60 KB for the tap gathers: 28 KB measured +3 %); instead
//   * normals, positions, previous-pixel positions, accept masks and sample counts of the NEXT tile are requested right
//     behind the phase barrier of this one — phase A is their last reader — and land during phase B;
//   * the albedo cells hold this tile's YCoCg values until phase B is over, so the next tile's albedo box is requested
//     behind the next barrier, on a barrier of its own, and is waited for where phase A first needs it: after the weighted
//     sum and the tap gathers of the thread's first pixel pair (AlbGate);
//   * the next tile's coefficients are loaded (into the other half of coef) between the phase barrier and phase B.
// Zone tiles of a strip stage their outgoing rows in the normal / position cells until they are pushed: for them the next
// tile's inputs are requested after the push.  Arithmetic, tile order of a CTA row and the zone bookkeeping are those of
// post_tma_kernel (same functions), so results are bit-identical to it.
// ------------------------------------------------------------------------------------------------
#ifndef BMFR_POST_PERSIST
#define BMFR_POST_PERSIST 0
#endif
#ifndef BMFR_POST_STAGGER_NS
#define BMFR_POST_STAGGER_NS 0
#endif
struct TileBoxes {  // 16-byte aligned box starts of a tile's six TMA copies (floor, also for negatives)
    int c_rgb, c_pp, c_u8, c1;
};
__device__ __forceinline__ void tile_geometry(const KParams& P, int bx, int by, TileGeom& G, TileBoxes& B) {
    G.x0 = bx * 32 - 16 + P.off_x;
    G.y0 = by * 32 - 16 + P.off_y;
    const int f_rgb = 3 * (G.x0 - 1), f_pp = 2 * (G.x0 - 1), b_u8 = G.x0 - 1;
    B.c_rgb = f_rgb & ~3; B.c_pp = f_pp & ~3; B.c_u8 = b_u8 & ~15;
    B.c1 = G.y0 - 1 - P.row0;
    G.sh_rgb = f_rgb - B.c_rgb; G.sh_pp = f_pp - B.c_pp; G.sh_u8 = b_u8 - B.c_u8;
}
// one thread: everything phase A reads except the albedo / the albedo.  Rows and columns outside the image (or the strip)
// arrive as zeros.
__device__ __forceinline__ void tile_request_inputs(PostStage& sh, const PostMaps& M, const TileBoxes& B) {
    mbar_expect_tx(&sh.bar, PostStage::TX - PT_HALO * PT_RGB_W * 4);
#if BMFR_L2_HINTS >= 2  // last readers of the reprojection's per-pixel outputs
    tma_load_tile_hint(&sh.pp[0][0], &M.pp, B.c_pp, B.c1, &sh.bar, BMFR_L2_ONCE);
    tma_load_tile_hint(&sh.acc[0][0], &M.accept, B.c_u8, B.c1, &sh.bar, BMFR_L2_ONCE);
#else
    tma_load_tile(&sh.pp[0][0], &M.pp, B.c_pp, B.c1, &sh.bar);
    tma_load_tile(&sh.acc[0][0], &M.accept, B.c_u8, B.c1, &sh.bar);
#endif
    tma_load_tile(&sh.spp[0][0], &M.spp, B.c_u8, B.c1, &sh.bar);
    tma_load_tile(&sh.nrm[0][0], &M.normals, B.c_rgb, B.c1, &sh.bar);
    tma_load_tile(&sh.pos[0][0], &M.positions, B.c_rgb, B.c1, &sh.bar);
}
__device__ __forceinline__ void tile_request_albedo(PostStage& sh, const PostMaps& M, const TileBoxes& B) {
    mbar_expect_tx(&sh.bar_alb, PT_HALO * PT_RGB_W * 4);
#if BMFR_L2_HINTS >= 2  // only reader of the albedo
    tma_load_tile_hint(&sh.alb[0][0], &M.albedo, B.c_rgb, B.c1, &sh.bar_alb, BMFR_L2_ONCE);
#else
    tma_load_tile(&sh.alb[0][0], &M.albedo, B.c_rgb, B.c1, &sh.bar_alb);
#endif
}

#if BMFR_POST_PERSIST
template <bool STRIP, int FS>
__global__ void __launch_bounds__(256, BMFR_POST_TMA_MIN_BLOCKS) post_persist_kernel(const __grid_constant__ KParams P, const __grid_constant__ PostMaps M) {
    constexpr int PX = 4, THREADS = 256, WARPS = 8;
    extern __shared__ __align__(128) unsigned char post_smem[];
    PostStage& sh = *reinterpret_cast<PostStage*>(post_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_x = P.blocks_x, nrows = P.by1 - P.by0, ntiles = tiles_x * nrows;
    if ((int)blockIdx.x >= ntiles) return;
    // tile (column bx, row i of the launch) -> block coordinates, geometry and box starts; a CTA's next tile is gridDim.x tiles
    // further (row-major): one division for the whole loop
    const int step_rows = (int)gridDim.x / tiles_x, step_cols = (int)gridDim.x - step_rows * tiles_x;
    auto locate = [&](int bx, int i, int& by, TileGeom& G, TileBoxes& B) {
        by = P.by0 + (STRIP ? halo_row_order(P.halo_p, i, nrows) : sweep_row(P, i, nrows));
        tile_geometry(P, bx, by, G, B);
    };
    auto advance = [&](int& bx, int& i) {
        bx += step_cols;
        i += step_rows;
        if (bx >= tiles_x) { bx -= tiles_x; ++i; }
    };
    int bx, ti;  // this CTA's current tile
    {
        ti = (int)blockIdx.x / tiles_x;
        bx = (int)blockIdx.x - ti * tiles_x;
        int by;
        TileGeom G;
        TileBoxes B;
        locate(bx, ti, by, G, B);
        if (tid == 0) {
            mbar_init(&sh.bar, 1);
            mbar_init(&sh.bar_alb, 1);
            mbar_fence_init();
            // Requested before the grid dependency is resolved: the caller's inputs and the reprojection's outputs are
            // complete by now (see post_tma_kernel).
            tile_request_inputs(sh, M, B);
            tile_request_albedo(sh, M, B);
        }
        pdl_wait();     // the fit of this frame is complete (weights, min/max)
        pdl_trigger();  // only now, so that "this frame's fit and reprojection are complete" also holds for the successor
        stamp_begin(P, 2);
#if BMFR_POST_STAGGER_NS > 0  // tuning: the CTAs of an SM (blockIdx.x = slot * SMs + SM, presumably) start their loops apart
        if (tid == 0) __nanosleep((blockIdx.x / (gridDim.x / BMFR_POST_TMA_MIN_BLOCKS)) * BMFR_POST_STAGGER_NS);
#endif
        load_coefficients_scaled<FS>(P, sh.coef[0], bx, by, warp, lane, WARPS);
    }
    for (int k = 0; ti < nrows; ++k) {
        int by;
        TileGeom G;
        TileBoxes B;
        locate(bx, ti, by, G, B);
        // strips: a tile near a strip edge waits for the neighbours' accumulated colour / TAA rows of the previous frame
        const bool zone = STRIP && halo_in_zone(P.halo_p, G.y0 - 1, G.y0 + PT_TILE + 1);
        if (zone) halo_poll(P.halo_p, halo_peek(P.halo_p, zone));
        __syncthreads();  // this tile's coefficients (and, k = 0, the barriers' initialisation) are visible; phase B of the previous tile is over
        if (k > 0 && tid == 0) tile_request_albedo(sh, M, B);
        mbar_wait_hot(&sh.bar, k & 1);
        f3 hist[PX];
        unsigned int live;
        const AlbGate gate{&sh.bar_alb, (unsigned int)(k & 1)};
        post_phase_a<STRIP, FS, PX, 3>(sh, P, G, gate, sh.coef[k & 1], tid, lane, warp, zone, hist, live);
        __syncthreads();  // the YCoCg cells are complete; nobody reads this tile's other inputs any more
        int bxn = bx, tin = ti;
        advance(bxn, tin);
        const bool more = tin < nrows;
        CoefRegs next_coef;
        if (more) {
            int byn;
            TileGeom Gn;
            TileBoxes Bn;
            locate(bxn, tin, byn, Gn, Bn);
            if (tid == 0 && !(STRIP && zone)) tile_request_inputs(sh, M, Bn);
            next_coef = fetch_coefficients<FS>(P, bxn, byn, warp, lane, WARPS);  // in flight across phase B
        }
        post_phase_b<STRIP, PX, 3>(sh, P, G, lane, warp, zone, hist, live);
        if (more) store_coefficients<FS>(next_coef, sh.coef[(k + 1) & 1], warp, lane, WARPS);
        if (STRIP && zone) {
            __syncthreads();  // the staged rows are complete
            post_push_rows<THREADS>(P, sh, G, tid);
            halo_finish(P.halo_p, halo_cta_pushes(P.halo_p, G.y0, G.y0 + PT_TILE));  // (starts with a barrier: the staged rows are read)
            if (more && tid == 0) {
                int byn;
                TileGeom Gn;
                TileBoxes Bn;
                locate(bxn, tin, byn, Gn, Bn);
                tile_request_inputs(sh, M, Bn);
            }
        }
        bx = bxn;
        ti = tin;
    }
    stamp_end(P, 2);
}
#endif  // BMFR_POST_PERSIST

#ifndef BMFR_POST_TMA
#define BMFR_POST_TMA 1
#endif

// Tensor maps of the frame's six read-once inputs, or false when the TMA path cannot be used (then post_kernel runs).
static bool post_maps(const KParams& P, PostMaps* M) {
    const int rows = P.row1 - P.row0;
    if (!BMFR_POST_TMA || (P.W & 15) != 0 || rows < PT_HALO || P.W * 3 < PT_RGB_W) return false;
    return bmfr_tensor_map_2d(P.cur_normals, 4, (long long)P.W * 3, rows, PT_RGB_W, PT_HALO, &M->normals) &&
           bmfr_tensor_map_2d(P.cur_positions, 4, (long long)P.W * 3, rows, PT_RGB_W, PT_HALO, &M->positions) &&
           bmfr_tensor_map_2d(P.albedo, 4, (long long)P.W * 3, rows, PT_RGB_W, PT_HALO, &M->albedo) &&
           bmfr_tensor_map_2d(P.prev_pixels, 4, (long long)P.W * 2, rows, PT_PP_W, PT_HALO, &M->pp) &&
           bmfr_tensor_map_2d(P.accept, 1, (long long)P.W, rows, PT_U8_W, PT_HALO, &M->accept) &&
           bmfr_tensor_map_2d(P.cur_spp, 1, (long long)P.W, rows, PT_U8_W, PT_HALO, &M->spp);
}

// Would launch_post() run the TMA-staged kernel for a whole image of this size?  (bmfr_create picks the history layout with it.)
bool post_uses_tma(int width, int rows) {
    return BMFR_POST_TMA && (width & 15) == 0 && rows >= PT_HALO && width * 3 >= PT_RGB_W && bmfr_encode_tiled_fn() != nullptr;
}

template <int FS>
static cudaError_t launch_post_fs(const KParams& P, cudaStream_t st) {
    const dim3 grid(P.blocks_x, P.by1 - P.by0);
    const bool strip = P.row0 != 0 || P.row1 != P.H;
    constexpr int PXW = BMFR_POST_PX, PXS = 4;  // pixels per thread: whole image / strips
    constexpr size_t smem_w = sizeof(PostStage) + BMFR_POST_SMEM_PAD;
    PostMaps M;
    if (post_maps(P, &M)) {
        static bool done[64] = {};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
        static int sms[64] = {};
        if (!done[dev]) {
            e = cudaFuncSetAttribute(post_tma_kernel<false, FS, PXW, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(post_tma_kernel<false, FS, PXW, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(post_tma_kernel<true, FS, PXS, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PostStage));
#if BMFR_POST_PERSIST
            if (e == cudaSuccess) e = cudaFuncSetAttribute(post_persist_kernel<false, FS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(post_persist_kernel<true, FS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PostStage));
#endif
            if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
            if (e != cudaSuccess) return e;
            done[dev] = true;
        }
        if (P.hist_stride == 4) {  // padded history: whole-image contexts only (bmfr_create decides)
            if (strip) return cudaErrorInvalidValue;
            return launch_pdl(!P.plain_launch, post_tma_kernel<false, FS, PXW, 4>, grid, dim3(1024 / PXW), smem_w, st, P, M);
        }
#if BMFR_POST_PERSIST
        {
            const int ntiles = P.blocks_x * (P.by1 - P.by0);
            const dim3 pgrid(ntiles < BMFR_POST_TMA_MIN_BLOCKS * sms[dev] ? ntiles : BMFR_POST_TMA_MIN_BLOCKS * sms[dev]);
            if (strip) return launch_pdl(!P.plain_launch, post_persist_kernel<true, FS>, pgrid, dim3(256), sizeof(PostStage), st, P, M);
            return launch_pdl(!P.plain_launch, post_persist_kernel<false, FS>, pgrid, dim3(256), smem_w, st, P, M);
        }
#endif
        if (strip) return launch_pdl(!P.plain_launch, post_tma_kernel<true, FS, PXS, 3>, grid, dim3(1024 / PXS), sizeof(PostStage), st, P, M);
        return launch_pdl(!P.plain_launch, post_tma_kernel<false, FS, PXW, 3>, grid, dim3(1024 / PXW), smem_w, st, P, M);
    }
    if (P.hist_stride != 3) return cudaErrorInvalidValue;  // the per-thread-load variant reads the reference's layout
    // widths that are no multiple of 16 (no tensor maps): the per-thread-load variant.  (Its 64+32-bit pixel accesses,
    // BMFR_POST_WIDE_ACCESS, cost 38 % more instructions for the same L1 wavefronts and stay a tuning switch.)
    const uintptr_t bits = (uintptr_t)P.cur_normals | (uintptr_t)P.cur_positions | (uintptr_t)P.albedo | (uintptr_t)P.accum_prev |
                           (uintptr_t)P.accum_cur | (uintptr_t)P.result_prev | (uintptr_t)P.result_cur | (uintptr_t)P.user_out;
    if constexpr (BMFR_POST_WIDE_ACCESS != 0) {
        if ((bits & 7) == 0) {
            if (strip) return launch_pdl(!P.plain_launch, post_kernel<true, true, FS>, grid, dim3(256), 0, st, P);
            return launch_pdl(!P.plain_launch, post_kernel<false, true, FS>, grid, dim3(256), 0, st, P);
        }
    }
    (void)bits;
    if (strip) return launch_pdl(!P.plain_launch, post_kernel<true, false, FS>, grid, dim3(256), 0, st, P);
    return launch_pdl(!P.plain_launch, post_kernel<false, false, FS>, grid, dim3(256), 0, st, P);
}

cudaError_t launch_post(const KParams& P, cudaStream_t st) {
    switch (P.feature_set) {
        case BMFR_FEATURE_SET_LINEAR: return launch_post_fs<BMFR_FEATURE_SET_LINEAR>(P, st);
        case BMFR_FEATURE_SET_POSITION: return launch_post_fs<BMFR_FEATURE_SET_POSITION>(P, st);
        default: return launch_post_fs<BMFR_FEATURE_SET_DEFAULT>(P, st);
    }
}
