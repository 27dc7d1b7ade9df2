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

// The same for the four pixels of a thread's column strip: every coefficient is fetched once and used four times.
template <int FS>
__device__ __forceinline__ void weighted_sum_px4(const f3 (&n)[4], const f3 (&p)[4], const float* __restrict__ cf, f3 (&out)[4]) {
    constexpr int F = FeatureSet<FS>::F;
    const float4* c4 = reinterpret_cast<const float4*>(cf);
    PostCoef<FS> pc;
    pc.load(cf);
    float ft[4][F - 1];
#pragma unroll
    for (int k = 0; k < 4; ++k) pc.features(n[k], p[k], ft[k]);
    const float4 w0 = c4[0];
    f3 acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = make_f3(w0.x, w0.y, w0.z);
#pragma unroll
    for (int f = 1; f < F; ++f) {
        const float4 w = c4[f];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc[k].x = fmaf(w.x, ft[k][f - 1], acc[k].x); acc[k].y = fmaf(w.y, ft[k][f - 1], acc[k].y); acc[k].z = fmaf(w.z, ft[k][f - 1], acc[k].z);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = clamp_negative(acc[k]);
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

__device__ __forceinline__ f3 to_ycocg(f3 c) {  // bmfr.cl:184-190
    return make_f3(c.x + 2.f * c.y + c.z, 2.f * c.x - 2.f * c.z, -c.x + 2.f * c.y - c.z);
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
// see the kernel), so with three CTAs per SM a tile's inputs land while its neighbours compute.  From shared
// memory a warp reads the same values at a stride of three words, which is bank-conflict free.
//   * Each pixel's tone-mapped YCoCg value is written over its own albedo cell (only that pixel's thread ever
//     reads the cell), so the neighbourhood planes of phase B need no memory of their own.
//   * What still goes through the L1 are the two 4-tap gathers.  A thread's vertically adjacent pixels share
//     their middle tap row whenever the reprojection is locally uniform (the common case; checked per pair), so
//     a pair loads three tap rows instead of four; all taps come from clamped addresses in one round of
//     independent loads, issued before the weighted sum so that its arithmetic overlaps their latency.
// Arithmetic and operation order per pixel are those of the per-thread-load variant above.
// ================================================================================================
#define PT_RGB_W 108  // floats per staged row of an interleaved-RGB image: 34 pixels + up to 3 floats of alignment shift
#define PT_PP_W 72    // floats per staged row of prev_pixels: 34 float2 + 2 floats of shift
#define PT_U8_W 64    // bytes per staged row of accept / spp: 34 + up to 15 bytes of shift

// A tile is 32 pixels wide and ROWS (32 or 16) high; the stage holds it with its one-pixel ring: HY = ROWS + 2 rows.
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
    float coef[9][PT_COEF];
    unsigned long long bar;
    static_assert(sizeof(float[HY_][PT_RGB_W]) % 128 == 96 && sizeof(float[HY_][PT_PP_W]) % 128 == 64 && (HY_ * PT_U8_W) % 128 == 0,
                  "TMA destinations must stay 128-byte aligned");
};

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
template <class SH>
__device__ __forceinline__ void put_cell_i(SH& sh, const TileGeom& G, int hx, int hy, f3 v) {
    float* p = &sh.alb[hy][G.sh_rgb + 3 * hx];
    p[0] = v.x; p[1] = v.y; p[2] = v.z;
}
// put_ycc() for the interleaved cells: the value of image pixel (x,y) also fills the out-of-image cells whose
// nearest in-image pixel it is (nobody reads an albedo there).
template <class SH>
__device__ __forceinline__ void put_ycc_i(SH& sh, const KParams& P, const TileGeom& G, int hx, int hy, int x, int y, f3 v) {
    put_cell_i(sh, G, hx, hy, v);
    const int ex = (x == 0) ? -1 : (x == P.W - 1) ? 1 : 0;
    const int ey = (y == 0) ? -1 : (y == P.H - 1) ? 1 : 0;
    if ((ex | ey) == 0) return;
    const bool okx = ex != 0 && (unsigned)(hx + ex) < PT_HALO, oky = ey != 0 && (unsigned)(hy + ey) < SH::HY;
    if (okx) put_cell_i(sh, G, hx + ex, hy, v);
    if (oky) put_cell_i(sh, G, hx, hy + ey, v);
    if (okx && oky) put_cell_i(sh, G, hx + ex, hy + ey, v);
}
template <class SH>
__device__ __forceinline__ PixelIn load_pixel_staged(const SH& sh, const KParams& P, const TileGeom& G, int hx, int hy, int x, int y) {
    PixelIn in;
    in.lp = pix_index(P, x, y);
    in.n = cell_f3(sh.nrm, G, hx, hy);
    in.p = cell_f3(sh.pos, G, hx, hy);
    in.alb = cell_f3(sh.alb, G, hx, hy);
    in.pp = *reinterpret_cast<const float2*>(&sh.pp[hy][G.sh_pp + 2 * hx]);
    in.accept = sh.acc[hy][G.sh_u8 + hx];
    in.spp = sh.spp[hy][G.sh_u8 + hx];
    return in;
}
// Zone CTAs of a strip (HaloK) do not send a mirrored row pixel by pixel: a pixel's accumulated colour goes into its own
// (already consumed) normal cell, its TAA result into its position cell, and after the last pixel the CTA sends whole row
// segments as 8-byte peer stores (post_push_rows) — 4-byte stores at a 12-byte stride make poor NVLink packets.
template <class SH>
__device__ __forceinline__ void stage_accum(SH& sh, const TileGeom& G, int hx, int hy, f3 v) {
    float* p = &sh.nrm[hy][G.sh_rgb + 3 * hx];
    p[0] = v.x; p[1] = v.y; p[2] = v.z;
}
template <class SH>
__device__ __forceinline__ void stage_result(SH& sh, const TileGeom& G, int hx, int hy, f3 v) {
    float* p = &sh.pos[hy][G.sh_rgb + 3 * hx];
    p[0] = v.x; p[1] = v.y; p[2] = v.z;
}
// All threads of a zone CTA, after a barrier: the tile's rows that a neighbour mirrors, both buffers.
template <class SH>
__device__ __forceinline__ void post_push_rows(const KParams& P, const SH& sh, const TileGeom& G, int tid) {
    constexpr int ROWS = SH::HY - 2, THREADS = 8 * ROWS;
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

// One pixel whose inputs are staged: ring pixels, and the pixels of a pair cut by a strip or image edge.
template <bool STRIP, int FS, class SH>
__device__ __forceinline__ bool staged_pixel(SH& sh, const KParams& P, const TileGeom& G, const float* cf, int hx, int hy, int x, int y,
                                             bool store, bool own, f3& hist, bool zone) {
    const PixelIn in = load_pixel_staged(sh, P, G, hx, hy, x, y);
    const f3 filtered = weighted_sum_px<FS>(in.n, in.p, cf);
    const bool temporal = own && history_sample<STRIP, false>(P, in.pp, hist);
    f3 accum;
    const f3 tone = accumulate_filtered_px<STRIP, false>(P, in.lp, filtered, in.accept, in.pp, in.spp, in.alb, store, x, y, false, &accum);
    if (STRIP && zone && store) stage_accum(sh, G, hx, hy, accum);
    put_ycc_i(sh, P, G, hx, hy, x, y, to_ycocg(tone));
    return temporal;
}

struct TapGeom {  // the 2x2 bilinear footprint of one pixel in the previous frame
    int pix, piy;
    float w[4];
};
__device__ __forceinline__ TapGeom tap_geom(float2 pp) {
    TapGeom t;
    t.pix = __float2int_rd(pp.x);
    t.piy = __float2int_rd(pp.y);
    const float frx = pp.x - (float)t.pix, fry = pp.y - (float)t.piy;
    const float omx = 1.f - frx, omy = 1.f - fry;
    t.w[0] = omx * omy; t.w[1] = frx * omy; t.w[2] = omx * fry; t.w[3] = frx * fry;
    return t;
}
// accumulate_filtered_data (bmfr.cl:778-849) and the TAA history sample (bmfr.cl:884-965) of one pixel from taps that
// are already in registers: a0 / r0 = the footprint's upper row (dx = 0, 1) of accumulated colour / TAA history, a1 / r1
// its lower row.  Same operation order as accumulate_filtered_px() / history_sample().
template <bool STRIP, class SH>
__device__ __forceinline__ bool resolve_pixel(SH& sh, const KParams& P, const TileGeom& G, const PixelIn& in, const TapGeom& t, f3 filtered,
                                              const f3 (&a0)[2], const f3 (&a1)[2], const f3 (&r0)[2], const f3 (&r1)[2], int hx, int hy, int x, int y,
                                              bool own, f3& hist, bool zone) {
    f3 prev = make_f3(0.f, 0.f, 0.f);
    float alpha = 1.f;
    // A strip that does not hold a wanted tap's row reports it (once, at the end: a store inside the loops would keep the
    // compiler from predicating the taps); the tap itself was fetched from a clamped address and the run is invalid anyway.
    const bool out0 = STRIP && (t.piy < P.state2_row0 || t.piy >= P.state2_row1);
    const bool out1 = STRIP && (t.piy + 1 < P.state2_row0 || t.piy + 1 >= P.state2_row1);
    bool missing = false;
    if (in.accept != 0) {
        float total = 0.f;
        missing = ((in.accept & 3u) != 0 && out0) || ((in.accept & 12u) != 0 && out1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (in.accept & (1u << i)) {  // taps are trusted, not re-checked: bmfr.cl:801-832
                const f3 pc = (i >> 1) ? a1[i & 1] : a0[i & 1];
                total += t.w[i];
                prev.x = fmaf(t.w[i], pc.x, prev.x);
                prev.y = fmaf(t.w[i], pc.y, prev.y);
                prev.z = fmaf(t.w[i], pc.z, prev.z);
            }
        }
        if (total > 0.f) {
            alpha = fmaxf(fast_rcp((float)in.spp), P.second_blend_alpha);  // bmfr.cl:838-839
            const float inv = fast_rcp(total);
            prev.x *= inv; prev.y *= inv; prev.z *= inv;
        }
    }
    const float oma = 1.f - alpha;
    const f3 accum = make_f3(fmaf(alpha, filtered.x, oma * prev.x), fmaf(alpha, filtered.y, oma * prev.y), fmaf(alpha, filtered.z, oma * prev.z));
    store_f3(P.accum_cur, in.lp, accum);
    if (STRIP && zone) stage_accum(sh, G, hx, hy, accum);
    const f3 tone = make_f3(tone_map_fast(in.alb.x * accum.x), tone_map_fast(in.alb.y * accum.y), tone_map_fast(in.alb.z * accum.z));
    put_ycc_i(sh, P, G, hx, hy, x, y, to_ycocg(tone));

    hist = make_f3(0.f, 0.f, 0.f);
    const bool temporal = own && !(t.pix < -1 || t.piy < -1 || t.pix >= P.W || t.piy >= P.H);  // bmfr.cl:884-890
    f3 hp = make_f3(0.f, 0.f, 0.f);
    float total = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bmfr.cl:929-960
        const int dx = i & 1, dy = i >> 1;
        const bool ok_y = dy ? (t.piy < P.H - 1) : (t.piy >= 0);
        const bool ok_x = dx ? (t.pix < P.W - 1) : (t.pix >= 0);
        missing = missing || (temporal && ok_x && ok_y && (dy ? out1 : out0));
        if (ok_x && ok_y) {
            const f3 pc = dy ? r1[dx] : r0[dx];
            hp.x = fmaf(t.w[i], pc.x, hp.x);
            hp.y = fmaf(t.w[i], pc.y, hp.y);
            hp.z = fmaf(t.w[i], pc.z, hp.z);
            total += t.w[i];
        }
    }
    if (STRIP && missing) *P.oob_flag = 1;
    if (!temporal) return false;
    const float inv = fast_rcp(total);  // 0 * inf = NaN on the image edge like the 0/0 of bmfr.cl:962
    hist = to_ycocg(make_f3(hp.x * inv, hp.y * inv, hp.z * inv));
    return true;
}

#ifndef BMFR_POST_TMA_MIN_BLOCKS
#define BMFR_POST_TMA_MIN_BLOCKS 3  // per SM for 32-row tiles; 16-row tiles: twice as many
#endif
#ifndef BMFR_POST_SMEM_PAD
#define BMFR_POST_SMEM_PAD 0  // extra dynamic shared memory per CTA of the whole-image instantiation: fewer CTAs per SM, more L1
#endif
#ifndef BMFR_POST_TILE_ROWS
// Tile height of the whole-image instantiation.  Measured at 1080p (profiles/r02 r3a): 16 (six 128-thread CTAs per SM, or
// five / four with more L1 through BMFR_POST_SMEM_PAD) 63.8 / 65.5 / 71.4 us against 57.7 us for 32 — 200 instead of 132 ring
// pixels per block and twice the per-CTA set-up cost more than the finer interleaving of the CTAs' waits gives.
#define BMFR_POST_TILE_ROWS 32
#endif
#ifndef BMFR_POST_WS4
#define BMFR_POST_WS4 0
#endif

// ROWS = 32 (default): one CTA of 256 threads per block, three per SM.  ROWS = 16 (tuning switch BMFR_POST_TILE_ROWS, slower):
// one CTA of 128 threads per half block, six per SM — the same 24 warps in six independent groups.
template <bool STRIP, int FS, int ROWS>
__global__ void __launch_bounds__(8 * ROWS, BMFR_POST_TMA_MIN_BLOCKS * 32 / ROWS) post_tma_kernel(const __grid_constant__ KParams P, const __grid_constant__ PostMaps M) {
    static_assert(ROWS == 32 || ROWS == 16, "a tile is a block or half a block");
    using Stage = PostStageT<ROWS + 2>;
    constexpr int HY = ROWS + 2, THREADS = 8 * ROWS;
    extern __shared__ __align__(128) unsigned char post_smem[];
    Stage& sh = *reinterpret_cast<Stage*>(post_smem);
    const int trow = STRIP ? halo_row_order(P.halo_p, blockIdx.y, gridDim.y) : sweep_row(P, blockIdx.y, gridDim.y);  // tile row of the launch
    const int half = ROWS == 32 ? 0 : (trow & 1);
    const int bx = blockIdx.x, by = P.by0 + (ROWS == 32 ? trow : (trow >> 1));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    TileGeom G;
    G.x0 = bx * 32 - 16 + P.off_x;
    G.y0 = by * 32 - 16 + P.off_y + ROWS * half;
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
        mbar_expect_tx(&sh.bar, Stage::TX);
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
    const bool zone = STRIP && halo_in_zone(P.halo_p, G.y0 - 1, G.y0 + ROWS + 1);
    const HaloPeek peek = halo_peek(P.halo_p, zone);
    pdl_wait();     // the fit of this frame is complete (weights, min/max)
    pdl_trigger();  // only now, so that "this frame's fit and reprojection are complete" also holds for the successor
    stamp_begin(P, 2);
    if (zone) halo_poll(P.halo_p, peek);

    load_coefficients<FS>(P, sh.coef, bx, by, warp, lane, THREADS / 32);
    __syncthreads();  // the coefficients and the barrier's initialisation are visible
    mbar_wait_hot(&sh.bar, 0);

    const int x = G.x0 + lane;
    const bool col_ok = x >= 0 && x < P.W;
    const int rlo = STRIP ? P.state2_row0 : 0, rhi = (STRIP ? P.state2_row1 : P.H) - 1;
    // phase A, interior: column strip x, rows 4*warp .. 4*warp+3, two vertically adjacent pixels at a time
    f3 hist[4];
    unsigned int live = 0;  // bit s: pixel s is written; bit 4+s: it takes the temporal path
#if BMFR_POST_WS4
    // the weighted sum of the whole strip first (shared-memory operands only), so that the coefficients are read once for
    // four pixels; strips cut by an edge take the pair path below
    f3 fl4[4];
    const bool strip_whole = col_ok && G.y0 + 4 * warp >= P.py0 && G.y0 + 4 * warp + 3 < P.py1;
    if (strip_whole) {
        f3 n4[4], p4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            n4[k] = cell_f3(sh.nrm, G, lane + 1, 4 * warp + k + 1);
            p4[k] = cell_f3(sh.pos, G, lane + 1, 4 * warp + k + 1);
        }
        weighted_sum_px4<FS>(n4, p4, sh.coef[4], fl4);
    }
#endif
#pragma unroll
    for (int s = 0; s < 4; s += 2) {
        const int ty = 4 * warp + s, y = G.y0 + ty;
        hist[s] = hist[s + 1] = make_f3(0.f, 0.f, 0.f);
        const bool v0 = col_ok && y >= P.py0 && y < P.py1, v1 = col_ok && y + 1 >= P.py0 && y + 1 < P.py1;
        const bool own0 = y >= P.own_y0 && y < P.own_y1, own1 = y + 1 >= P.own_y0 && y + 1 < P.own_y1;
        if (v0 && v1) {
            const PixelIn i0 = load_pixel_staged(sh, P, G, lane + 1, ty + 1, x, y), i1 = load_pixel_staged(sh, P, G, lane + 1, ty + 2, x, y + 1);
            bool t0 = false, t1 = false;
            if (P.frame > 0) {
                const TapGeom g0 = tap_geom(i0.pp), g1 = tap_geom(i1.pp);
                // one round of independent loads from clamped addresses: tap rows R0, R1 of the upper pixel, R2 of the lower;
                // the lower pixel's upper row is R1 when the two footprints are stacked (checked below)
                const int cx0[2] = {min(max(g0.pix, 0), P.W - 1), min(max(g0.pix + 1, 0), P.W - 1)};
                const int cx1[2] = {min(max(g1.pix, 0), P.W - 1), min(max(g1.pix + 1, 0), P.W - 1)};
                const int ry0[2] = {min(max(g0.piy, rlo), rhi), min(max(g0.piy + 1, rlo), rhi)};
                const int ry1[2] = {min(max(g1.piy, rlo), rhi), min(max(g1.piy + 1, rlo), rhi)};
                f3 A[3][2], R[3][2];
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const unsigned int l0 = pix_index(P, cx0[dx], ry0[0]), l1 = pix_index(P, cx0[dx], ry0[1]), l2 = pix_index(P, cx1[dx], ry1[1]);
                    A[0][dx] = load_f3(P.accum_prev, l0); A[1][dx] = load_f3(P.accum_prev, l1); A[2][dx] = load_f3(P.accum_prev, l2);
                    R[0][dx] = load_f3(P.result_prev, l0); R[1][dx] = load_f3(P.result_prev, l1); R[2][dx] = load_f3(P.result_prev, l2);
                }
                f3 fl0, fl1;
#if BMFR_POST_WS4
                if (strip_whole) { fl0 = fl4[s]; fl1 = fl4[s + 1]; }
                else
#endif
                weighted_sum_px2<FS>(i0.n, i0.p, i1.n, i1.p, sh.coef[4], fl0, fl1);  // its arithmetic overlaps the gathers
                t0 = resolve_pixel<STRIP>(sh, P, G, i0, g0, fl0, A[0], A[1], R[0], R[1], lane + 1, ty + 1, x, y, own0, hist[s], zone);
                if (!(cx1[0] == cx0[0] && cx1[1] == cx0[1] && ry1[0] == ry0[1])) {  // footprints not stacked (rare): fetch the row
#pragma unroll
                    for (int dx = 0; dx < 2; ++dx) {
                        const unsigned int l = pix_index(P, cx1[dx], ry1[0]);
                        A[1][dx] = load_f3(P.accum_prev, l);
                        R[1][dx] = load_f3(P.result_prev, l);
                    }
                }
                t1 = resolve_pixel<STRIP>(sh, P, G, i1, g1, fl1, A[1], A[2], R[1], R[2], lane + 1, ty + 2, x, y + 1, own1, hist[s + 1], zone);
            } else {  // frame 0: no temporal path, alpha = 1 (bmfr.cl:784, 884)
                f3 fl0, fl1;
#if BMFR_POST_WS4
                if (strip_whole) { fl0 = fl4[s]; fl1 = fl4[s + 1]; }
                else
#endif
                weighted_sum_px2<FS>(i0.n, i0.p, i1.n, i1.p, sh.coef[4], fl0, fl1);
                store_f3(P.accum_cur, i0.lp, fl0);
                store_f3(P.accum_cur, i1.lp, fl1);
                if (STRIP && zone) {
                    stage_accum(sh, G, lane + 1, ty + 1, fl0);
                    stage_accum(sh, G, lane + 1, ty + 2, fl1);
                }
                put_ycc_i(sh, P, G, lane + 1, ty + 1, x, y,
                          to_ycocg(make_f3(tone_map_fast(i0.alb.x * fl0.x), tone_map_fast(i0.alb.y * fl0.y), tone_map_fast(i0.alb.z * fl0.z))));
                put_ycc_i(sh, P, G, lane + 1, ty + 2, x, y + 1,
                          to_ycocg(make_f3(tone_map_fast(i1.alb.x * fl1.x), tone_map_fast(i1.alb.y * fl1.y), tone_map_fast(i1.alb.z * fl1.z))));
            }
            live |= ((own0 ? 1u : 0u) | (t0 ? 16u : 0u)) << s;
            live |= ((own1 ? 1u : 0u) | (t1 ? 16u : 0u)) << (s + 1);
        } else if (v0) {  // a strip or image edge cuts the pair
            const bool t = staged_pixel<STRIP, FS>(sh, P, G, sh.coef[4], lane + 1, ty + 1, x, y, true, own0, hist[s], zone);
            live |= ((own0 ? 1u : 0u) | (t ? 16u : 0u)) << s;
        } else if (v1) {
            const bool t = staged_pixel<STRIP, FS>(sh, P, G, sh.coef[4], lane + 1, ty + 2, x, y + 1, true, own1, hist[s + 1], zone);
            live |= ((own1 ? 1u : 0u) | (t ? 16u : 0u)) << (s + 1);
        }
    }
    // phase A, ring: 2 * 33 + 2 * (HY - 1) pixels (132 / 100), coefficients of the pixel's own block — for a half tile the
    // row below the upper half / above the lower half belongs to the same block
    if (tid < 2 * (PT_HALO - 1) + 2 * (HY - 1)) {
        int hx, hy;
        if (tid < PT_HALO - 1) { hx = tid; hy = 0; }
        else if (tid < PT_HALO - 1 + HY - 1) { hx = PT_HALO - 1; hy = tid - (PT_HALO - 1); }
        else if (tid < 2 * (PT_HALO - 1) + HY - 1) { hx = PT_HALO - 1 - (tid - (PT_HALO - 1 + HY - 1)); hy = HY - 1; }
        else { hx = 0; hy = HY - 1 - (tid - (2 * (PT_HALO - 1) + HY - 1)); }
        const int rx = G.x0 + hx - 1, ry = G.y0 + hy - 1;
        if (rx >= 0 && rx < P.W && ry >= P.py0 && ry < P.py1) {
            const int nrow = (hy == 0) ? ((ROWS == 32 || half == 0) ? 0 : 3) : (hy == HY - 1) ? ((ROWS == 32 || half == 1) ? 6 : 3) : 3;
            const int nb = nrow + ((hx == 0) ? 0 : (hx == PT_HALO - 1) ? 2 : 1);
            f3 unused;
            staged_pixel<STRIP, FS>(sh, P, G, sh.coef[nb], hx, hy, rx, ry, false, false, unused, zone);
        }
    }
    __syncthreads();

    // phase B: clamp the history samples to the neighbourhood box, component by component (bmfr.cl:893-920, 967-969).  Halo
    // rows 4*warp .. 4*warp+5 cover the 3x3 neighbourhoods of the strip; this thread's column is lane+1.
    f3 mine[4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float ctr[6], rmin[6], rmax[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const float* row = &sh.alb[4 * warp + r][G.sh_rgb + c];
            const float l = row[3 * lane], m = row[3 * lane + 3], rr = row[3 * lane + 6];
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
    // blend and store (bmfr.cl:971-973)
    const float a = P.taa_blend_alpha, oma = 1.f - P.taa_blend_alpha;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (!(live & (1u << s))) continue;
        const unsigned int lp = pix_index(P, x, G.y0 + 4 * warp + s);
        f3 out = from_ycocg(mine[s]);  // this pixel's tone-mapped colour
        if (live & (16u << s)) {
            const f3 pr = from_ycocg(hist[s]);
            out = make_f3(fmaf(a, out.x, oma * pr.x), fmaf(a, out.y, oma * pr.y), fmaf(a, out.z, oma * pr.z));
        }
        store_f3(P.result_cur, lp, out);
        if (P.user_out) store_f3(P.user_out, lp, out);
        if (STRIP && zone) stage_result(sh, G, lane + 1, 4 * warp + s + 1, out);
    }
    if (STRIP && zone) {
        __syncthreads();  // the staged rows are complete
        post_push_rows(P, sh, G, tid);
        halo_finish(P.halo_p, halo_cta_pushes(P.halo_p, G.y0, G.y0 + ROWS));
    }
    stamp_end(P, 2);
}

#ifndef BMFR_POST_TMA
#define BMFR_POST_TMA 1
#endif

// Tensor maps of the frame's six read-once inputs, or false when the TMA path cannot be used (then post_kernel runs).
static bool post_maps(const KParams& P, PostMaps* M, int box_rows) {
    const int rows = P.row1 - P.row0;
    if (!BMFR_POST_TMA || (P.W & 15) != 0 || rows < PT_HALO || P.W * 3 < PT_RGB_W) return false;
    return bmfr_tensor_map_2d(P.cur_normals, 4, (long long)P.W * 3, rows, PT_RGB_W, box_rows, &M->normals) &&
           bmfr_tensor_map_2d(P.cur_positions, 4, (long long)P.W * 3, rows, PT_RGB_W, box_rows, &M->positions) &&
           bmfr_tensor_map_2d(P.albedo, 4, (long long)P.W * 3, rows, PT_RGB_W, box_rows, &M->albedo) &&
           bmfr_tensor_map_2d(P.prev_pixels, 4, (long long)P.W * 2, rows, PT_PP_W, box_rows, &M->pp) &&
           bmfr_tensor_map_2d(P.accept, 1, (long long)P.W, rows, PT_U8_W, box_rows, &M->accept) &&
           bmfr_tensor_map_2d(P.cur_spp, 1, (long long)P.W, rows, PT_U8_W, box_rows, &M->spp);
}

template <int FS>
static cudaError_t launch_post_fs(const KParams& P, cudaStream_t st) {
    const dim3 grid(P.blocks_x, P.by1 - P.by0);
    const bool strip = P.row0 != 0 || P.row1 != P.H;
    // tile height: whole images run half tiles (six 128-thread CTAs per SM), strips whole blocks — their zone bookkeeping
    // (fill_halo counts zone CTAs per block row) and row staging are laid out for 32 x 32 tiles
    constexpr int WHOLE_ROWS = BMFR_POST_TILE_ROWS;
    using StageW = PostStageT<WHOLE_ROWS + 2>;
    using StageS = PostStageT<32 + 2>;
    constexpr size_t smem_w = sizeof(StageW) + BMFR_POST_SMEM_PAD;
    PostMaps M;
    if (post_maps(P, &M, (strip ? 32 : WHOLE_ROWS) + 2)) {
        static bool done[64] = {};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
        if (!done[dev]) {
            e = cudaFuncSetAttribute(post_tma_kernel<false, FS, WHOLE_ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(post_tma_kernel<true, FS, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StageS));
            if (e != cudaSuccess) return e;
            done[dev] = true;
        }
        if (strip) return launch_pdl(!P.plain_launch, post_tma_kernel<true, FS, 32>, grid, dim3(256), sizeof(StageS), st, P, M);
        const dim3 grid_w(P.blocks_x, (P.by1 - P.by0) * (32 / WHOLE_ROWS));
        return launch_pdl(!P.plain_launch, post_tma_kernel<false, FS, WHOLE_ROWS>, grid_w, dim3(8 * WHOLE_ROWS), smem_w, st, P, M);
    }
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
