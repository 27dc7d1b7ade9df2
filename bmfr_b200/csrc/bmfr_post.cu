// FUSED mode, second half of the frame: weighted_sum + accumulate_filtered_data + taa
// (bmfr.cl:703-758, 761-857, 860-974) in one pass; `filtered` and `tone_mapped` never reach HBM.
//
// One CTA per 32x32 tile of the frame's shifted block grid, so the block's 42 fit coefficients are
// CTA-uniform and sit in registers.  Thread (lane, warp) owns the column strip x = x0 + lane,
// rows y0 + 4*warp .. +3: every global access of a warp is 32 consecutive pixels of one row (the
// interleaved-RGB stride of 12 B keeps each 32-bit load on 3-4 cache lines), and the four 3x3 TAA
// neighbourhoods of a strip share their row minima / maxima.
//   phase A : filtered -> accumulated -> tone-mapped for the tile and a one-pixel ring (ring pixels
//             use their own block's coefficients), written to shared memory as YCoCg planes;
//             neighbours outside the image are filled with the nearest in-image pixel, which leaves
//             the min / max over the in-image neighbours unchanged (bmfr.cl:900-920) and removes
//             every per-neighbour test.
//   phase B : taa for the tile interior from shared memory.
// Nothing here is compared bitwise with the reference (the inputs already carry the fit's
// rounding), so this translation unit is compiled with FMA contraction and uses the fast
// reciprocal; tests/test_gpu_parity.py holds it to the 1e-3 / 60 dB colour tolerance.
#include "bmfr_kernels.h"

#include "bmfr_device.cuh"

#define PT_TILE 32
#define PT_HALO (PT_TILE + 2)
#define PT_STRIDE 36  // floats per shared-memory row (>= 34)

struct PostShared {
    float ycc[3][PT_HALO][PT_STRIDE];
    float coef[BMFR_FEATURES * 3 + BMFR_FEATURES_SCALED * 2 + 2];
};

__device__ __forceinline__ float fast_rcp(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ float tone_map_fast(float v) {  // clamp(powr(max(0,v), 0.454545f), 0, 1), bmfr.cl:852-856
    v = fmaxf(0.f, v);
    v = exp2f(0.454545f * __log2f(v));
    return __saturatef(v);
}

// weighted_sum for one pixel, bmfr.cl:725-750
__device__ __forceinline__ f3 weighted_sum_px(f3 n, f3 p, const float* __restrict__ w, const float* __restrict__ mi) {
    const float feat[BMFR_FEATURES] = {1.f,
                                       n.x,
                                       n.y,
                                       n.z,
                                       (p.x - mi[0]) * mi[1],
                                       (p.y - mi[2]) * mi[3],
                                       (p.z - mi[4]) * mi[5],
                                       (p.x * p.x - mi[6]) * mi[7],
                                       (p.y * p.y - mi[8]) * mi[9],
                                       (p.z * p.z - mi[10]) * mi[11]};
    f3 c = make_f3(w[0], w[1], w[2]);
#pragma unroll
    for (int f = 1; f < BMFR_FEATURES; ++f) {
        c.x = fmaf(w[f * 3 + 0], feat[f], c.x);
        c.y = fmaf(w[f * 3 + 1], feat[f], c.y);
        c.z = fmaf(w[f * 3 + 2], feat[f], c.z);
    }
    c.x = c.x < 0.f ? 0.f : c.x;  // keeps NaN like the reference, bmfr.cl:750
    c.y = c.y < 0.f ? 0.f : c.y;
    c.z = c.z < 0.f ? 0.f : c.z;
    return c;
}

// accumulate_filtered_data for one pixel, bmfr.cl:778-856.  Returns the tone-mapped colour.
template <bool STRIP>
__device__ __forceinline__ f3 accumulate_filtered_px(const KParams& P, unsigned int lp, f3 filtered, float2 pp,
                                                     unsigned int accept, bool store) {
    f3 prev = make_f3(0.f, 0.f, 0.f);
    float alpha = 1.f;
    if (P.frame > 0 && accept != 0) {
        const int pix = __float2int_rd(pp.x), piy = __float2int_rd(pp.y);
        const float frx = pp.x - (float)pix, fry = pp.y - (float)piy;
        const float omx = 1.f - frx, omy = 1.f - fry;
        const float w[4] = {omx * omy, frx * omy, omx * fry, frx * fry};
        float total = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (accept & (1u << i)) {  // taps are trusted, not re-checked: bmfr.cl:801-832
                const int sx = pix + (i & 1), sy = piy + (i >> 1);
                if (STRIP && (sy < P.row0 || sy >= P.row1)) {
                    *P.oob_flag = 1;
                    continue;
                }
                const f3 pc = load_f3(P.accum_prev, pix_index(P, sx, sy));
                total += w[i];
                prev.x = fmaf(w[i], pc.x, prev.x);
                prev.y = fmaf(w[i], pc.y, prev.y);
                prev.z = fmaf(w[i], pc.z, prev.z);
            }
        }
        if (total > 0.f) {
            alpha = fmaxf(fast_rcp((float)P.cur_spp[lp]), P.second_blend_alpha);  // bmfr.cl:838-839
            const float inv = fast_rcp(total);
            prev.x *= inv;
            prev.y *= inv;
            prev.z *= inv;
        }
    }
    const float oma = 1.f - alpha;
    const f3 accum = make_f3(fmaf(alpha, filtered.x, oma * prev.x), fmaf(alpha, filtered.y, oma * prev.y),
                             fmaf(alpha, filtered.z, oma * prev.z));
    if (store) store_f3(P.accum_cur, lp, accum);
    const f3 alb = load_f3(P.albedo, lp);
    return make_f3(tone_map_fast(alb.x * accum.x), tone_map_fast(alb.y * accum.y), tone_map_fast(alb.z * accum.z));
}

__device__ __forceinline__ f3 to_ycocg(f3 c) {  // bmfr.cl:184-190
    return make_f3(c.x + 2.f * c.y + c.z, 2.f * c.x - 2.f * c.z, -c.x + 2.f * c.y - c.z);
}
__device__ __forceinline__ f3 from_ycocg(f3 c) {  // bmfr.cl:192-198
    return make_f3(0.25f * (c.x + c.y - c.z), 0.25f * (c.x + c.z), 0.25f * (c.x - c.y - c.z));
}

// Stores the YCoCg value of image pixel (x,y) at halo cell (hx,hy) and replicates it into the
// out-of-image cells whose nearest in-image pixel it is.
__device__ __forceinline__ void put_cell(PostShared& sh, int hx, int hy, f3 v) {
    sh.ycc[0][hy][hx] = v.x;
    sh.ycc[1][hy][hx] = v.y;
    sh.ycc[2][hy][hx] = v.z;
}
__device__ __forceinline__ void put_ycc(PostShared& sh, const KParams& P, int hx, int hy, int x, int y, f3 v) {
    put_cell(sh, hx, hy, v);
    const int ex = (x == 0) ? -1 : (x == P.W - 1) ? 1 : 0;
    const int ey = (y == 0) ? -1 : (y == P.H - 1) ? 1 : 0;
    const bool okx = ex != 0 && (unsigned)(hx + ex) < PT_HALO, oky = ey != 0 && (unsigned)(hy + ey) < PT_HALO;
    if (okx) put_cell(sh, hx + ex, hy, v);
    if (oky) put_cell(sh, hx, hy + ey, v);
    if (okx && oky) put_cell(sh, hx + ex, hy + ey, v);
}

template <bool STRIP>
__global__ void __launch_bounds__(256, 2) post_kernel(const __grid_constant__ KParams P) {
    __shared__ __align__(16) PostShared sh;
    const int bx = blockIdx.x, by = P.by0 + blockIdx.y;
    const int group = by * P.blocks_x + bx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = bx * 32 - 16 + P.off_x, y0 = by * 32 - 16 + P.off_y;  // tile origin in image coordinates
    constexpr int NW = BMFR_FEATURES * 3, NM = BMFR_FEATURES_SCALED * 2;

    if (tid < NW) sh.coef[tid] = __ldg(P.weights + (size_t)group * NW + tid);
    else if (tid < NW + NM) sh.coef[tid] = __ldg(P.mins_inv + (size_t)group * NM + tid - NW);
    __syncthreads();

    const int x = x0 + lane;
    const bool col_ok = x >= 0 && x < P.W;
    f3 mine[4];
    float2 pp[4];
    bool have[4];
    {
        float w[NW], mi[NM];
#pragma unroll
        for (int i = 0; i < NW; ++i) w[i] = sh.coef[i];
#pragma unroll
        for (int i = 0; i < NM; ++i) mi[i] = sh.coef[NW + i];
        // phase A, interior
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int ty = 4 * warp + s, y = y0 + ty;
            have[s] = col_ok && y >= P.py0 && y < P.py1;
            if (have[s]) {
                const unsigned int lp = pix_index(P, x, y);
                const f3 filtered = weighted_sum_px(load_f3(P.cur_normals, lp), load_f3(P.cur_positions, lp), w, mi);
                pp[s] = __ldg(P.prev_pixels + lp);
                mine[s] = accumulate_filtered_px<STRIP>(P, lp, filtered, pp[s], __ldg(P.accept + lp), true);
                put_ycc(sh, P, lane + 1, ty + 1, x, y, to_ycocg(mine[s]));
            }
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
            const int g = k3_group(P, rx, ry);
            const unsigned int lp = pix_index(P, rx, ry);
            float w[NW], mi[NM];
#pragma unroll
            for (int i = 0; i < NW; ++i) w[i] = __ldg(P.weights + (size_t)g * NW + i);
#pragma unroll
            for (int i = 0; i < NM; ++i) mi[i] = __ldg(P.mins_inv + (size_t)g * NM + i);
            const f3 filtered = weighted_sum_px(load_f3(P.cur_normals, lp), load_f3(P.cur_positions, lp), w, mi);
            const f3 tone = accumulate_filtered_px<STRIP>(P, lp, filtered, __ldg(P.prev_pixels + lp), __ldg(P.accept + lp), false);
            put_ycc(sh, P, hx, hy, rx, ry, to_ycocg(tone));
        }
    }
    __syncthreads();

    // phase B: clamp bounds of the strip's four pixels, plane by plane.  Halo rows 4*warp .. 4*warp+5
    // cover the 3x3 neighbourhoods of tile rows 4*warp .. 4*warp+3; this thread's column is lane+1.
    float lo[4][3], hi[4][3];
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
            lo[s][c] = (min_box + min_cross) * 0.5f;  // bmfr.cl:967-968
            hi[s][c] = (max_box + max_cross) * 0.5f;
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int y = y0 + 4 * warp + s;
        if (!have[s] || y < P.own_y0 || y >= P.own_y1) continue;
        const unsigned int lp = pix_index(P, x, y);
        const f3 my_new = mine[s];
        const int pix = __float2int_rd(pp[s].x), piy = __float2int_rd(pp[s].y);
        f3 out = my_new;
        if (!(P.frame == 0 || pix < -1 || piy < -1 || pix >= P.W || piy >= P.H)) {  // bmfr.cl:884-890
            const float frx = pp[s].x - (float)pix, fry = pp[s].y - (float)piy;
            const float omx = 1.f - frx, omy = 1.f - fry;
            const float w[4] = {omx * omy, frx * omy, omx * fry, frx * fry};
            f3 prev = make_f3(0.f, 0.f, 0.f);
            float total = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {  // bmfr.cl:929-960
                const int dx = i & 1, dy = i >> 1;
                const bool ok_y = dy ? (piy < P.H - 1) : (piy >= 0);
                const bool ok_x = dx ? (pix < P.W - 1) : (pix >= 0);
                if (ok_x && ok_y) {
                    const int sy = piy + dy;
                    if (STRIP && (sy < P.row0 || sy >= P.row1)) {
                        *P.oob_flag = 1;
                        continue;
                    }
                    const f3 pc = load_f3(P.result_prev, pix_index(P, pix + dx, sy));
                    prev.x = fmaf(w[i], pc.x, prev.x);
                    prev.y = fmaf(w[i], pc.y, prev.y);
                    prev.z = fmaf(w[i], pc.z, prev.z);
                    total += w[i];
                }
            }
            const float inv = 1.0f / total;  // 0/0 on the image edge like bmfr.cl:962
            const f3 py = to_ycocg(make_f3(prev.x * inv, prev.y * inv, prev.z * inv));
            const f3 cl = make_f3(fminf(fmaxf(py.x, lo[s][0]), hi[s][0]), fminf(fmaxf(py.y, lo[s][1]), hi[s][1]),
                                  fminf(fmaxf(py.z, lo[s][2]), hi[s][2]));
            const f3 pr = from_ycocg(cl);
            const float a = P.taa_blend_alpha, oma = 1.f - P.taa_blend_alpha;
            out = make_f3(fmaf(a, my_new.x, oma * pr.x), fmaf(a, my_new.y, oma * pr.y), fmaf(a, my_new.z, oma * pr.z));
        }
        store_f3(P.result_cur, lp, out);
        if (P.user_out) store_f3(P.user_out, lp, out);
    }
}

cudaError_t launch_post(const KParams& P, cudaStream_t st) {
    const dim3 grid(P.blocks_x, P.by1 - P.by0);
    if (P.row0 != 0 || P.row1 != P.H) post_kernel<true><<<grid, 256, 0, st>>>(P);
    else post_kernel<false><<<grid, 256, 0, st>>>(P);
    return cudaGetLastError();
}
