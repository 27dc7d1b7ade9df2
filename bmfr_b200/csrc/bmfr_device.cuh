// Device-side building blocks of the BMFR hot path (sm_100a).
//
// Arithmetic convention (DESIGN.md "Arithmetic"): this translation unit is compiled with
// --fmad=false, so every a*b+c written with operators is two correctly rounded IEEE operations,
// evaluated left to right exactly like the CPU oracle (gcc -ffp-contract=off).  That is what makes
// the reprojection outputs (accept mask, spp, floor(prev pixel)) and every other per-pixel stage
// bit-identical to the oracle on the same inputs.  Fused multiply-adds are used only where written
// explicitly as fmaf(): inside the least-squares fit, whose results are compared within tolerance.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/bmfr_b200.h"
#include "bmfr_kernels.h"

// Programmatic dependent launch (the FUSED kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization): a kernel may start while its predecessor in the
// stream is still draining; pdl_wait() blocks until the predecessor has completed and its writes are
// visible, pdl_trigger() lets the successor's CTAs be scheduled as soon as SM resources free up.  Both
// are no-ops for a normal launch.  Every kernel of the chain calls pdl_wait() before it completes, so
// "predecessor complete" is transitive along the frame loop.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Launch with programmatic stream serialization (host side, all three FUSED kernels); pdl = false: a plain
// launch (the kernel's griddepcontrol instructions are then no-ops) for the overlapped-frames mode, where
// the three kernels live on three streams and are ordered by events.
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// ---------------------------------------------------------------------------------------------
// Feature lists (bmfr_params.feature_set).  The reference pastes a compile-time string into its kernels
// (NOT_SCALED_FEATURE_BUFFERS / SCALED_FEATURE_BUFFERS, bmfr.cpp:63-77 -> FEATURE_BUFFERS, bmfr.cl:447-453,724-729); here
// the FUSED fit and post pass are instantiated for a fixed set of lists.  F = features (the constant 1 first), NSC = the
// scaled ones (at the end), then three colour columns: NCOL non-constant columns per block matrix.
// ---------------------------------------------------------------------------------------------
template <int FS> struct FeatureSet;
template <> struct FeatureSet<BMFR_FEATURE_SET_DEFAULT> {  // 1, n.xyz | p.xyz, p.xyz^2 (bmfr.cpp:65-77)
    static constexpr int F = 10, NSC = 6;
    static constexpr bool NORMALS = true, SQUARES = true;
};
template <> struct FeatureSet<BMFR_FEATURE_SET_LINEAR> {   // 1, n.xyz | p.xyz
    static constexpr int F = 7, NSC = 3;
    static constexpr bool NORMALS = true, SQUARES = false;
};
template <> struct FeatureSet<BMFR_FEATURE_SET_POSITION> { // 1 | p.xyz, p.xyz^2
    static constexpr int F = 7, NSC = 6;
    static constexpr bool NORMALS = false, SQUARES = true;
};
// the non-constant columns of one matrix row from the pixel's normal, world position and accumulated colour
template <int FS>
__device__ __forceinline__ void feature_columns(float* __restrict__ dst, const float* __restrict__ n, const float* __restrict__ p, const float* __restrict__ col) {
    int k = 0;
    if (FeatureSet<FS>::NORMALS) { dst[k] = n[0]; dst[k + 1] = n[1]; dst[k + 2] = n[2]; k += 3; }
    dst[k] = p[0]; dst[k + 1] = p[1]; dst[k + 2] = p[2]; k += 3;
    if (FeatureSet<FS>::SQUARES) { dst[k] = p[0] * p[0]; dst[k + 1] = p[1] * p[1]; dst[k + 2] = p[2] * p[2]; k += 3; }
    dst[k] = col[0]; dst[k + 1] = col[1]; dst[k + 2] = col[2];
}

// ---------------------------------------------------------------------------------------------
// In-kernel halo exchange of strip contexts (HaloK in bmfr_kernels.h).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// profile = 2 (bmfr_get_fused_kernel_busy_ms): when did the first CTA of kernel k start, when did the last one end?
__device__ __forceinline__ void stamp_begin(const KParams& P, int k) {
    if (P.stamps != nullptr && threadIdx.x == 0 && threadIdx.y == 0) atomicMin(P.stamps + 2 * k, globaltimer_ns());
}
__device__ __forceinline__ void stamp_end(const KParams& P, int k) {
    if (P.stamps != nullptr && threadIdx.x == 0 && threadIdx.y == 0) atomicMin(P.stamps + 2 * k + 1, ~globaltimer_ns());
}

// Sweep direction of the whole-image kernels.  The three kernels of a frame read the same 75 MB (normals, world positions,
// accumulated colour) one after the other and the L2 holds the part touched last, so each kernel starts where its
// predecessor finished: reprojection and post pass run down the image, the fit up (BMFR_ZIGZAG: and the next frame the other
// way round, because the post pass of frame f ends where the reprojection of frame f + 1 finds its history).
#ifndef BMFR_ZIGZAG
#define BMFR_ZIGZAG 0
#endif
__device__ __forceinline__ bool sweep_down(int frame) { return !BMFR_ZIGZAG || (frame & 1) == 0; }
__device__ __forceinline__ int sweep_row(const KParams& P, int i, int n) { return sweep_down(P.frame) ? i : n - 1 - i; }

// Zone rows early and spread out: CTA row i of n (blockIdx.y, which the hardware hands out in ascending order) -> the row
// it works on.  The Z zone rows take every k-th position from the start (k = n / Z), the interior rows fill the rest: the
// zone CTAs' peer stores then drain over NVLink while interior CTAs compute (all zone CTAs at once would hold most of
// the SM slots while their stores and system fences complete), and the neighbours still see the flag before this
// kernel ends.  Without neighbours (rows_top = rows_bot = 0) this is the identity.
__device__ __forceinline__ int halo_row_order(const HaloK& h, int i, int n) {
    const int top = h.rows_top, Z = h.rows_top + h.rows_bot;
    if (Z <= 0 || Z >= n) return i;
    const int k = n / Z;
    const int placed = min(Z, (i + k - 1) / k);  // zone rows at positions before i
    if (i % k == 0 && i / k < Z) {
        const int j = i / k;                      // the j-th zone row: top ones first, then the bottom ones
        return j < top ? j : n - h.rows_bot + (j - top);
    }
    return top + (i - placed);                    // the (i - placed)-th interior row
}
#ifndef BMFR_HALO_RELEASE_GPU
#define BMFR_HALO_RELEASE_GPU 1
#endif
#ifndef BMFR_HALO_ACQUIRE_FENCE
#define BMFR_HALO_ACQUIRE_FENCE 0
#endif
// Does the CTA that covers image rows [ya, yb) belong to the zone?
__device__ __forceinline__ bool halo_in_zone(const HaloK& h, int ya, int yb) { return h.active && (ya < h.zone_y[0] || yb > h.zone_y[1]); }
// Prologue of a zone CTA (all threads call it): wait for the neighbours' rows.  Bounded: after timeout_ns the context is
// marked failed (flags[2], sticky) and the kernel goes on with whatever it finds; a failed context never signals.
//
// No acquire fence follows the poll.  A fence.acq_rel (what __threadfence_system() is) ends in CCTL.IVALL: it drops the
// whole L1 of the SM, under every co-resident CTA's tap gathers — with 40 % of a strip's CTAs in the zone that cost 19 us
// of the reprojection's 118 and 25 us of the post pass's 144 at 8K over eight GPUs (profiles/r02_n8_timeline.md).  What
// the fence would protect against is a stale L1 copy of a halo row.  There is none: the L1 is invalidated when a grid
// starts, and within a grid nothing reads a neighbour-written row before this poll has seen the flag — interior CTAs
// never touch such rows (that is what makes a CTA a zone CTA), zone CTAs poll first, and the fit and the TMA-staged
// inputs of the post pass only read rows this context wrote itself.  The neighbour's rows are in this GPU's L2 (the
// point of coherence for peer stores) before its flag is: it raises the flag behind a system-scope release.  The poll is
// a volatile (strong, system-scope) load, the barrier orders the CTA's later loads after it, and an SM does not issue a
// load ahead of the branch that depends on the polled value.
// The flags only grow, so a zone CTA reads them once at its very start (halo_peek, thread 0: four independent loads whose L2
// round trip hides behind the CTA's set-up and its wait for the predecessor grid) and polls only if that look was too early.
struct HaloPeek {
    unsigned int e0, e1, l0, l1;
};
__device__ __forceinline__ HaloPeek halo_peek(const HaloK& h, bool zone) {
    HaloPeek k{0u, 0u, 0u, 0u};
    if (zone && threadIdx.x == 0 && threadIdx.y == 0) {
        volatile unsigned int* f = h.flags;
        k.e0 = f[0]; k.e1 = f[1]; k.l0 = f[4]; k.l1 = f[5];
    }
    return k;
}
__device__ __forceinline__ void halo_poll(const HaloK& h, HaloPeek k) {
#ifdef BMFR_DEBUG_NO_POLL  // timing experiments only: results are wrong
    return;
#endif
    if (threadIdx.x == 0 && threadIdx.y == 0 && (h.wait_early | h.wait_late) != 0) {
        volatile unsigned int* f = h.flags;
        unsigned long long t0 = 0;
        for (bool first = true;; first = false) {
            unsigned int e0 = k.e0, e1 = k.e1, l0 = k.l0, l1 = k.l1;
            if (!first) { e0 = f[0]; e1 = f[1]; l0 = f[4]; l1 = f[5]; }  // four independent loads: one L2 round trip
            const bool a = !h.side_on[0] || (e0 >= h.wait_early && l0 >= h.wait_late);
            const bool b = !h.side_on[1] || (e1 >= h.wait_early && l1 >= h.wait_late);
            if (a && b) break;
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > h.timeout_ns) {
                f[2] = 1;
                break;
            }
            __nanosleep(100);
        }
#if BMFR_HALO_ACQUIRE_FENCE
        __threadfence_system();
#endif
    }
    __syncthreads();
}
// Epilogue of a zone CTA (all threads call it, after their last store): the last zone CTA of the launch raises the flags.
// A CTA that pushed rows counts itself with a system-scope RELEASE (its peer stores, ordered before thread 0 by the
// barrier, are performed before the count; unlike a fence this carries no L1 invalidation); the last one to arrive reads the
// count at the end of that release sequence, fences once (acquire + release) and stores the flags.
__device__ __forceinline__ void halo_finish(const HaloK& h, bool pushed) {
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        unsigned int done;
#if BMFR_HALO_RELEASE_GPU
        if (pushed) asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(done) : "l"(h.done_counter) : "memory");
#else
        if (pushed) asm volatile("atom.add.release.sys.global.u32 %0, [%1], 1;" : "=r"(done) : "l"(h.done_counter) : "memory");
#endif
        else done = atomicAdd(h.done_counter, 1u);  // only read halo rows: nothing to publish
        if (done + 1 == h.zone_ctas) {
            __threadfence_system();
            *h.done_counter = 0;  // for the next launch (ordered after this one by the stream)
            if (*(volatile unsigned int*)(h.flags + 2) == 0) {
#pragma unroll
                for (int s = 0; s < 2; ++s)
                    if (h.side_on[s]) *(volatile unsigned int*)h.peer_flag[s] = h.signal_value;
                __threadfence_system();
            }
        }
    }
}
// Does a CTA that covers image rows [ya, yb) store any row a neighbour mirrors?
__device__ __forceinline__ bool halo_cta_pushes(const HaloK& h, int ya, int yb) {
#ifdef BMFR_DEBUG_NO_PUSH  // timing experiments only: results are wrong
    return false;
#endif
    return (h.side_on[0] && ya < h.push_y1[0] && yb > h.push_y0[0]) || (h.side_on[1] && ya < h.push_y1[1] && yb > h.push_y0[1]);
}
// Index of image pixel (x, y) in the neighbour's buffers on side s, or -1 when the row is not mirrored there.
__device__ __forceinline__ long long halo_peer_index(const HaloK& h, const KParams& P, int s, int x, int y) {
    return (h.side_on[s] && y >= h.push_y0[s] && y < h.push_y1[s]) ? (long long)(y - h.peer_row0[s]) * P.W + x : -1;
}

#ifndef BMFR_STREAM_LOADS
#define BMFR_STREAM_LOADS 0
#endif

struct f3 {
    float x, y, z;
};

__device__ __forceinline__ f3 make_f3(float x, float y, float z) { return f3{x, y, z}; }
// Pixel indices are 32-bit (an 8K strip has 33 M pixels, 100 M floats); one widening per access.
__device__ __forceinline__ f3 load_f3(const float* __restrict__ b, unsigned int i) {
    const float* p = b + (size_t)(i * 3u);
    return f3{__ldg(p), __ldg(p + 1), __ldg(p + 2)};
}
// Read-once (streaming) data: do not allocate the line in L1, so that the L1 keeps the lines the 4-tap
// gathers share between neighbouring pixels.
__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// L2 eviction priorities (tuning switch BMFR_L2_HINTS): a frame's normals, world positions and
// accumulated colour are read by all three kernels within ~100 us — "keep" asks the L2 to hold on to them; inputs that are
// read exactly once are marked "once".  The operand is the descriptor createpolicy.fractional.L2::evict_* ... 1.0 produces.
#ifndef BMFR_L2_HINTS
#define BMFR_L2_HINTS 2
#endif
#define BMFR_L2_KEEP 0x14F0000000000000ull  // evict_last
#define BMFR_L2_ONCE 0x12F0000000000000ull  // evict_first
__device__ __forceinline__ float ldg_hint(const float* p, unsigned long long policy) {
    float v;
    asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(policy));  // read-only data: free to move
    return v;
}
__device__ __forceinline__ void stg_hint(float* p, float v, unsigned long long policy) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ f3 load_f3_hint(const float* __restrict__ b, unsigned int i, unsigned long long policy) {
    const float* p = b + (size_t)(i * 3u);
    return f3{ldg_hint(p, policy), ldg_hint(p + 1, policy), ldg_hint(p + 2, policy)};
}
__device__ __forceinline__ void store_f3_hint(float* __restrict__ b, unsigned int i, f3 v, unsigned long long policy) {
    float* p = b + (size_t)(i * 3u);
    stg_hint(p, v.x, policy);
    stg_hint(p + 1, v.y, policy);
    stg_hint(p + 2, v.z, policy);
}
__device__ __forceinline__ f3 load_f3_stream(const float* __restrict__ b, unsigned int i) {
#if BMFR_STREAM_LOADS
    const float* p = b + (size_t)(i * 3u);
    return f3{ldg_stream(p), ldg_stream(p + 1), ldg_stream(p + 2)};
#else
    return load_f3(b, i);
#endif
}
__device__ __forceinline__ void store_f3(float* __restrict__ b, unsigned int i, f3 v) {
    float* p = b + (size_t)(i * 3u);
    p[0] = v.x;
    p[1] = v.y;
    p[2] = v.z;
}
// The same 12 bytes with two memory instructions instead of three.  Pixel i starts at byte 12 i, which
// is 8-byte aligned for even i: an even pixel is (64-bit, 32-bit), an odd one (32-bit, 64-bit).  Both
// shapes are issued as one 64-bit access at the aligned half and one 32-bit access at the other, and
// three selects put the components back in order.  A warp-wide access touches the same cache lines
// either way, so this removes a third of the L1 wavefronts of every interleaved-RGB access.
__device__ __forceinline__ f3 load_f3_wide(const float* __restrict__ b, unsigned int i) {
    const bool odd = (i & 1u) != 0;
    const float* p = b + (size_t)(i * 3u);
    const float2 w = __ldg(reinterpret_cast<const float2*>(p + (odd ? 1 : 0)));
    const float s = __ldg(p + (odd ? 0 : 2));
    return f3{odd ? s : w.x, odd ? w.x : w.y, odd ? w.y : s};
}
__device__ __forceinline__ void store_f3_wide(float* __restrict__ b, unsigned int i, f3 v) {
    const bool odd = (i & 1u) != 0;
    float* p = b + (size_t)(i * 3u);
    *reinterpret_cast<float2*>(p + (odd ? 1 : 0)) = odd ? make_float2(v.y, v.z) : make_float2(v.x, v.y);
    p[odd ? 0 : 2] = odd ? v.x : v.z;
}
__device__ __forceinline__ float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ f3 sub3(f3 a, f3 b) { return f3{a.x - b.x, a.y - b.y, a.z - b.z}; }

// Packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2): one issue slot for two IEEE-rounded operations.  Each
// half is the same round-to-nearest operation as the scalar instruction.  CAUTION for code that must
// match the oracle bit for bit: ptxas 12.9 contracts mul.rn.f32x2 followed by add/sub.rn.f32x2 into
// FFMA2 even with --fmad=false (it honours .rn only for scalars; checked with cuobjdump), so a product
// that feeds a sum must be a scalar __fmul_rn; packed sums of scalar products are not contracted.
__device__ __forceinline__ unsigned long long f2_bits(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 bits_f2(unsigned long long b) { return *reinterpret_cast<float2*>(&b); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
    return bits_f2(d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(d);
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return bits_f2(d);
}
__device__ __forceinline__ float2 dup2(float v) { return make_float2(v, v); }

__device__ __forceinline__ unsigned int pix_index(const KParams& P, int x, int y) {
    return (unsigned int)((y - P.row0) * P.W + x);
}

// mirror(), bmfr.cl:209-216
__device__ __forceinline__ int mirror_index(int i, int size) {
    if (i < 0) return -i - 1;
    if (i >= size) return 2 * size - i - 1;
    return i;
}

// scale(), bmfr.cl:200-205: (v - min) / (max - min) if |max - min| > 1 else v - min.
// The block-uniform divisor is inverted once (scale_factor) and applied as a multiply in both the
// fit and the weighted sum, so the two stay consistent; the scaled value differs from the
// reference's quotient by at most one ulp (inside the colour tolerance; mins_maxs stay exact).
__device__ __forceinline__ float scale_factor(float mn, float mx) {
    const float d = mx - mn;
    return (fabsf(d) > 1.0f) ? 1.0f / d : 1.0f;
}
__device__ __forceinline__ float scale_feature(float v, float mn, float inv) { return (v - mn) * inv; }

// random(), bmfr.cl:162-171 — integer hash; float(a) / float(UINT_MAX) == float(a) * 2^-32 exactly
__host__ __device__ __forceinline__ float bmfr_random(unsigned int a) {
    a = (a + 0x7ed55d16u) + (a << 12);
    a = (a ^ 0xc761c23cu) ^ (a >> 19);
    a = (a + 0x165667b1u) + (a << 5);
    a = (a + 0xd3a2646cu) ^ (a << 9);
    a = (a + 0xfd7046c5u) + (a << 3);
    a = (a ^ 0xb55a4f09u) ^ (a >> 16);
    return (float)a / 4294967296.0f;
}

// ---------------------------------------------------------------------------------------------
// K1 per work-item: accumulate_noisy_data, bmfr.cl:319-453, for the (already mirrored) pixel x,y.
// ---------------------------------------------------------------------------------------------
struct K1Pixel {
    f3 normal, position, new_color;
    float prev_x, prev_y;
    unsigned char accept, spp;
};

// STRIP = the context holds only a band of rows: gathers are checked against it (oob_flag).
// wp: the pixel's world position, already loaded (callers that walk over several pixels fetch the next
// pixel's position before they start this one, so that the two dependent memory round trips of a pixel
// — position -> reprojection -> taps — overlap across pixels).
// n, cur: the pixel's shading normal and this frame's noisy colour, already loaded too.
// carry: a thread that walks down a column hands the lower tap row of one pixel to the next; when the reprojection is
// locally uniform (the common case, checked per pixel) that row IS the upper tap row of the pixel below, and its ten
// values need not be fetched again.  The same values from the same addresses: nothing changes in the arithmetic.
struct K1Carry {
    int cx0, cx1, ry;  // clamped coordinates the row was fetched from; ry < 0: nothing carried
    f3 tp[2], tn[2], tc[2];
    float ts[2];
};
template <bool STRIP>
__device__ __forceinline__ K1Pixel k1_pixel_core(const KParams& P, int x, int y, f3 wp, f3 n, f3 cur, K1Carry* carry = nullptr) {
    K1Pixel r;
    float pfx = (float)x, pfy = (float)y;  // bmfr.cl:325
    unsigned int accept = 0;
    float blend_alpha = 1.f;
    f3 prev = make_f3(0.f, 0.f, 0.f);
    float sample_spp = 0.f;

    if (P.frame > 0) {
        const float* M = P.cam;
        // dot(M.s048c, (p,1)) etc., left to right; bmfr.cl:343-349.  The sums of x and y share packed
        // additions; the products stay scalar (see the caution above): every operation is the IEEE
        // operation of the reference expression, in the same order.
        float2 cxy = fadd2(make_float2(__fmul_rn(M[0], wp.x), __fmul_rn(M[1], wp.x)),
                           make_float2(__fmul_rn(M[4], wp.y), __fmul_rn(M[5], wp.y)));
        cxy = fadd2(cxy, make_float2(__fmul_rn(M[8], wp.z), __fmul_rn(M[9], wp.z)));
        cxy = fadd2(cxy, make_float2(M[12], M[13]));  // M * 1.f
        const float cw = ((M[3] * wp.x + M[7] * wp.y) + M[11] * wp.z) + M[15] * 1.f;
        cxy = make_float2(cxy.x / cw, cxy.y / cw);
        cxy = fmul2(fadd2(cxy, dup2(1.f)), dup2(0.5f));  // (c + 1) / 2: a division by 2 is the exact product with 0.5
        const float2 pf = fsub2(make_float2(__fmul_rn(cxy.x, (float)P.W), __fmul_rn(cxy.y, (float)P.H)),
                                make_float2(P.poff_x, P.poff_y1));  // bmfr.cl:352-355
        pfx = pf.x;
        pfy = pf.y;
        const int pix = __float2int_rd(pfx);  // convert_int2_rtn, bmfr.cl:356
        const int piy = __float2int_rd(pfy);
        const float2 fr = fsub2(pf, make_float2((float)pix, (float)piy));
        const float2 om = fsub2(dup2(1.f), fr);
        const float w[4] = {om.x * om.y, fr.x * om.y, om.x * fr.y, fr.x * fr.y};  // bmfr.cl:367-370
        float total = 0.f;
        // The four taps are fetched unconditionally from clamped addresses (one round of independent
        // loads instead of three dependent ones) and the reference's nested tests (bmfr.cl:380-404)
        // become predicates; a rejected tap contributes nothing, exactly as in the branchy form.
        const int rlo = P.row0, rhi = P.row1 - 1;
        f3 tp[4], tn[4], tc[4];
        float ts[4];
        bool valid[4];
        const int cx[2] = {min(max(pix, 0), P.W - 1), min(max(pix + 1, 0), P.W - 1)};
        const int ry[2] = {min(max(piy, rlo), rhi), min(max(piy + 1, rlo), rhi)};
        const bool reuse = carry != nullptr && carry->ry == ry[0] && carry->cx0 == cx[0] && carry->cx1 == cx[1];
        bool missing = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int sx = pix + (i & 1), sy = piy + (i >> 1);
            const bool inimg = sx >= 0 && sy >= 0 && sx < P.W && sy < P.H;  // bmfr.cl:380-381
            const bool held = !STRIP || (sy >= P.row0 && sy < P.row1);
            missing = missing || (inimg && !held);  // strip + halo does not hold this row: reported once, below
            valid[i] = inimg && held;
            if (i < 2 && reuse) {  // the row the pixel above left behind
                tp[i] = carry->tp[i]; tn[i] = carry->tn[i]; tc[i] = carry->tc[i]; ts[i] = carry->ts[i];
            } else {
                const unsigned int ls = pix_index(P, cx[i & 1], ry[i >> 1]);
                tp[i] = load_f3(P.prev_positions, ls);
                tn[i] = load_f3(P.prev_normals, ls);
                tc[i] = load_f3(P.prev_noisy_acc, ls);
                ts[i] = (float)__ldg(P.prev_spp + ls);
            }
        }
        if (STRIP && missing) *P.oob_flag = 1;
        if (carry != nullptr) {
            carry->cx0 = cx[0]; carry->cx1 = cx[1]; carry->ry = ry[1];
#pragma unroll
            for (int i = 0; i < 2; ++i) { carry->tp[i] = tp[2 + i]; carry->tn[i] = tn[2 + i]; carry->tc[i] = tc[2 + i]; carry->ts[i] = ts[2 + i]; }
        }
        // position and normal differences ride in the two halves of a pair (bmfr.cl:388-404);
        // (sample_spp, prev.x) and (prev.y, prev.z) are accumulated as pairs (bmfr.cl:407-415)
        const float2 refx = make_float2(wp.x, n.x), refy = make_float2(wp.y, n.y), refz = make_float2(wp.z, n.z);
        float2 acc_sx = make_float2(0.f, 0.f), acc_yz = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 dx = fsub2(make_float2(tp[i].x, tn[i].x), refx);
            const float2 dy = fsub2(make_float2(tp[i].y, tn[i].y), refy);
            const float2 dz = fsub2(make_float2(tp[i].z, tn[i].z), refz);
            const float2 d2 = fadd2(fadd2(make_float2(__fmul_rn(dx.x, dx.x), __fmul_rn(dx.y, dx.y)),
                                          make_float2(__fmul_rn(dy.x, dy.x), __fmul_rn(dy.y, dy.y))),
                                    make_float2(__fmul_rn(dz.x, dz.x), __fmul_rn(dz.y, dz.y)));  // (|dp|^2, |dn|^2), dot left to right
            const bool ok = valid[i] && (d2.x < P.pos_limit) && (d2.y < P.nrm_limit);
            if (ok) {
                accept |= 1u << i;
                acc_sx = fadd2(acc_sx, make_float2(__fmul_rn(w[i], ts[i]), __fmul_rn(w[i], tc[i].x)));
                acc_yz = fadd2(acc_yz, make_float2(__fmul_rn(w[i], tc[i].y), __fmul_rn(w[i], tc[i].z)));
                total = total + w[i];
            }
        }
        sample_spp = acc_sx.x;
        prev = make_f3(acc_sx.y, acc_yz.x, acc_yz.y);
        if (total > 0.f) {  // bmfr.cl:421-429
            prev.x = prev.x / total;
            prev.y = prev.y / total;
            prev.z = prev.z / total;
            sample_spp = sample_spp / total;
            blend_alpha = 1.f / (sample_spp + 1.f);
            blend_alpha = fmaxf(blend_alpha, P.blend_alpha);
        }
    }
    unsigned int new_spp = 1;  // bmfr.cl:433-441
    if (blend_alpha < 1.f) {
        if (sample_spp > 254.f) {
            new_spp = 255;
        } else {
            int s = __float2int_rn(sample_spp);  // convert_uchar_sat_rte
            s = min(max(s, 0), 255);
            new_spp = (unsigned int)(s + 1) & 0xFFu;
        }
    }
    const float oma = 1.f - blend_alpha;
    const float2 nxy = fadd2(make_float2(__fmul_rn(blend_alpha, cur.x), __fmul_rn(blend_alpha, cur.y)),
                             make_float2(__fmul_rn(oma, prev.x), __fmul_rn(oma, prev.y)));
    r.new_color = make_f3(nxy.x, nxy.y, blend_alpha * cur.z + oma * prev.z);  // bmfr.cl:444-445
    r.normal = n;
    r.position = wp;
    r.prev_x = pfx;
    r.prev_y = pfy;
    r.accept = (unsigned char)accept;
    r.spp = (unsigned char)new_spp;
    return r;
}

// The 13 values K1 stores per work-item (bmfr.cl:448-476): features with NaN replaced by 0.  A NaN
// position gives a NaN square, so scrubbing the position first yields the same 13 values.
__device__ __forceinline__ float scrub_nan(float v) { return isnan(v) ? 0.f : v; }
__device__ __forceinline__ void k1_features(const K1Pixel& r, float* f) {
    const float px = scrub_nan(r.position.x), py = scrub_nan(r.position.y), pz = scrub_nan(r.position.z);
    f[0] = 1.f;
    f[1] = scrub_nan(r.normal.x);
    f[2] = scrub_nan(r.normal.y);
    f[3] = scrub_nan(r.normal.z);
    f[4] = px;
    f[5] = py;
    f[6] = pz;
    f[7] = px * px;
    f[8] = py * py;
    f[9] = pz * pz;
    f[10] = scrub_nan(r.new_color.x);
    f[11] = scrub_nan(r.new_color.y);
    f[12] = scrub_nan(r.new_color.z);
}

template <bool STRIP>
__device__ __forceinline__ K1Pixel k1_pixel(const KParams& P, int x, int y, f3 wp) {
    const unsigned int lp = pix_index(P, x, y);
    return k1_pixel_core<STRIP>(P, x, y, wp, load_f3_stream(P.cur_normals, lp), load_f3_stream(P.cur_noisy, lp));
}
template <bool STRIP>
__device__ __forceinline__ K1Pixel k1_pixel(const KParams& P, int x, int y) {
    return k1_pixel<STRIP>(P, x, y, load_f3(P.cur_positions, pix_index(P, x, y)));
}

// ---------------------------------------------------------------------------------------------
// K3 per pixel: weighted_sum, bmfr.cl:725-750.  w = 30 weights, mi = 6 (min, 1/range) pairs of the
// block.  The sum is compared with the reference within tolerance, so it is written with fmaf().
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ f3 k3_pixel(f3 n, f3 p, const float* __restrict__ w, const float* __restrict__ mi) {
    float feat[BMFR_FEATURES] = {1.f, n.x, n.y, n.z, p.x, p.y, p.z, p.x * p.x, p.y * p.y, p.z * p.z};
    f3 c = make_f3(0.f, 0.f, 0.f);
#pragma unroll
    for (int f = 0; f < BMFR_FEATURES; ++f) {
        float v = feat[f];
        if (f >= BMFR_FEATURES_NOT_SCALED)
            v = scale_feature(v, mi[(f - BMFR_FEATURES_NOT_SCALED) * 2], mi[(f - BMFR_FEATURES_NOT_SCALED) * 2 + 1]);
        c.x = fmaf(w[f * 3 + 0], v, c.x);
        c.y = fmaf(w[f * 3 + 1], v, c.y);
        c.z = fmaf(w[f * 3 + 2], v, c.z);
    }
    c.x = c.x < 0.f ? 0.f : c.x;  // keeps NaN like the reference, bmfr.cl:750
    c.y = c.y < 0.f ? 0.f : c.y;
    c.z = c.z < 0.f ? 0.f : c.z;
    return c;
}

// group_index of weighted_sum, bmfr.cl:718-722
__device__ __forceinline__ int k3_group(const KParams& P, int x, int y) {
    return ((x + 16 - P.off_x) >> 5) + ((y + 16 - P.off_y) >> 5) * P.blocks_x;
}

// ---------------------------------------------------------------------------------------------
// K4 per pixel: accumulate_filtered_data, bmfr.cl:778-856.
// ---------------------------------------------------------------------------------------------
#ifndef BMFR_FAST_POW
#define BMFR_FAST_POW 1
#endif
__device__ __forceinline__ float tone_map(float v) {  // clamp(powr(max(0,v), 0.454545f), 0, 1)
    v = fmaxf(0.f, v);
#if BMFR_FAST_POW
    // powr through the SFU: |error| of __log2f is ~2^-22 absolute, so the result is within a few
    // 1e-7 relative of powf on [0, 2^10] — three orders below the colour tolerance
    v = exp2f(0.454545f * __log2f(v));
#else
    v = powf(v, 0.454545f);
#endif
    return fminf(fmaxf(v, 0.f), 1.f);
}

__device__ __forceinline__ void k4_pixel(const KParams& P, unsigned int lp, f3 filtered, float prev_x, float prev_y,
                                         unsigned int accept, f3& accum, f3& tone) {
    f3 prev = make_f3(0.f, 0.f, 0.f);
    float blend_alpha = 1.f;
    if (P.frame > 0 && accept > 0) {
        const int pix = __float2int_rd(prev_x), piy = __float2int_rd(prev_y);
        const float frx = prev_x - (float)pix, fry = prev_y - (float)piy;
        const float omx = 1.f - frx, omy = 1.f - fry;
        const float w[4] = {omx * omy, frx * omy, omx * fry, frx * fry};
        float total = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (accept & (1u << i)) {  // taps are trusted, not re-checked: bmfr.cl:801-832
                const int sx = pix + (i & 1), sy = piy + (i >> 1);
                if (sy < P.state2_row0 || sy >= P.state2_row1) {
                    *P.oob_flag = 1;
                    continue;
                }
                total = total + w[i];
                const f3 pc = load_f3(P.accum_prev, pix_index(P, sx, sy));
                prev.x = prev.x + w[i] * pc.x;
                prev.y = prev.y + w[i] * pc.y;
                prev.z = prev.z + w[i] * pc.z;
            }
        }
        if (total > 0.f) {
            blend_alpha = 1.f / (float)P.cur_spp[lp];  // bmfr.cl:838-839
            blend_alpha = fmaxf(blend_alpha, P.second_blend_alpha);
            prev.x = prev.x / total;
            prev.y = prev.y / total;
            prev.z = prev.z / total;
        }
    }
    const float oma = 1.f - blend_alpha;
    accum = make_f3(blend_alpha * filtered.x + oma * prev.x, blend_alpha * filtered.y + oma * prev.y,
                    blend_alpha * filtered.z + oma * prev.z);
    const f3 alb = load_f3(P.albedo, lp);
    tone = make_f3(tone_map(alb.x * accum.x), tone_map(alb.y * accum.y), tone_map(alb.z * accum.z));
}

// ---------------------------------------------------------------------------------------------
// K5 helpers: taa, bmfr.cl:184-198,893-973.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ f3 rgb_to_ycocg(f3 c) {
    return make_f3((c.x * 1.f + c.y * 2.f) + c.z * 1.f, (c.x * 2.f + c.y * 0.f) + c.z * -2.f,
                   (c.x * -1.f + c.y * 2.f) + c.z * -1.f);
}
__device__ __forceinline__ f3 ycocg_to_rgb(f3 c) {
    return make_f3((c.x * 0.25f + c.y * 0.25f) + c.z * -0.25f, (c.x * 0.25f + c.y * 0.f) + c.z * 0.25f,
                   (c.x * 0.25f + c.y * -0.25f) + c.z * -0.25f);
}
__device__ __forceinline__ f3 min3(f3 a, f3 b) { return make_f3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
__device__ __forceinline__ f3 max3(f3 a, f3 b) { return make_f3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }

struct TaaBox {
    f3 min_box, min_cross, max_box, max_cross;
};
__device__ __forceinline__ void taa_box_init(TaaBox& b) {
    b.min_box = b.min_cross = make_f3(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
    b.max_box = b.max_cross = make_f3(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
}
__device__ __forceinline__ void taa_box_add_ycocg(TaaBox& b, f3 s, bool cross) {
    if (cross) {
        b.min_cross = min3(b.min_cross, s);
        b.max_cross = max3(b.max_cross, s);
    }
    b.min_box = min3(b.min_box, s);
    b.max_box = max3(b.max_box, s);
}
__device__ __forceinline__ void taa_box_add(TaaBox& b, f3 rgb, bool cross) { taa_box_add_ycocg(b, rgb_to_ycocg(rgb), cross); }

// Everything of taa after the 3x3 neighbourhood: bilinear history fetch, clamp, blend (bmfr.cl:922-973).
__device__ __forceinline__ f3 taa_resolve(const KParams& P, f3 my_new, const TaaBox& b, float prev_x, float prev_y,
                                          int pix, int piy) {
    f3 prev = make_f3(0.f, 0.f, 0.f);
    float total = 0.f;
    const float frx = prev_x - (float)pix, fry = prev_y - (float)piy;
    const float omx = 1.f - frx, omy = 1.f - fry;
    const float w[4] = {omx * omy, frx * omy, omx * fry, frx * fry};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int dx = i & 1, dy = i >> 1;
        const bool ok_y = dy ? (piy < P.H - 1) : (piy >= 0);
        const bool ok_x = dx ? (pix < P.W - 1) : (pix >= 0);
        if (ok_x && ok_y) {
            const int sy = piy + dy;
            if (sy < P.state2_row0 || sy >= P.state2_row1) {
                *P.oob_flag = 1;
                continue;
            }
            const f3 pc = load_f3(P.result_prev, pix_index(P, pix + dx, sy));
            prev.x = prev.x + w[i] * pc.x;
            prev.y = prev.y + w[i] * pc.y;
            prev.z = prev.z + w[i] * pc.z;
            total = total + w[i];
        }
    }
    prev.x = prev.x / total;  // may be 0/0 on the image edge, as in bmfr.cl:962
    prev.y = prev.y / total;
    prev.z = prev.z / total;
    const f3 py = rgb_to_ycocg(prev);
    const f3 mn = make_f3((b.min_box.x + b.min_cross.x) / 2.f, (b.min_box.y + b.min_cross.y) / 2.f,
                          (b.min_box.z + b.min_cross.z) / 2.f);
    const f3 mx = make_f3((b.max_box.x + b.max_cross.x) / 2.f, (b.max_box.y + b.max_cross.y) / 2.f,
                          (b.max_box.z + b.max_cross.z) / 2.f);
    // clamp(x, lo, hi) = fmin(fmax(x, lo), hi)
    const f3 cl = make_f3(fminf(fmaxf(py.x, mn.x), mx.x), fminf(fmaxf(py.y, mn.y), mx.y), fminf(fmaxf(py.z, mn.z), mx.z));
    const f3 pr = ycocg_to_rgb(cl);
    const float a = P.taa_blend_alpha, oma = 1.f - P.taa_blend_alpha;
    return make_f3(a * my_new.x + oma * pr.x, a * my_new.y + oma * pr.y, a * my_new.z + oma * pr.z);
}

// ---------------------------------------------------------------------------------------------
// The fit: fitter, bmfr.cl:490-700, for one 32x32 block held in registers.
//
// The reference runs a Householder QR of the 1024x13 matrix [features | colour] with 577 work-group
// barriers and 193 passes over global memory per block, then back-substitutes R(0..9,0..9) x =
// R(0..9,10..12).  Here the matrix never leaves registers: thread `tid` owns rows tid + 256*s (the
// reference's IN_ACCESS ownership, bmfr.cl:90-97) and R is computed by a two-level TSQR — each warp
// factors its own 128 rows without any block barrier, then one warp factors the eight stacked 10x13
// triangles and back-substitutes.
//
// Each reflector is taken against a virtual zero pivot row appended to the matrix.  With pivot
// entry 0 the reference's formulas (bmfr.cl:580-587,650) collapse: |x| = sqrt(S_k), u_k = -|x|,
// u_length_squared = 2 S_k, and a_j -= a_k * S_j / S_k with S_j = a_k . a_j, R_kj = S_j / sqrt(S_k)
// — i.e. one modified-Gram-Schmidt step on the augmented matrix, which is the Householder QR of
// [0; A] (Bjorck) and backward stable for the least-squares solution.  It needs no pivot broadcast
// and no per-lane masks.  R of a QR with positive diagonal is unique, so this is the reference's
// R(0..9, 0..12) up to fp32 rounding (tests: colour within 1e-3 relative of the oracle).
//
// The per-reflector reduction of the 13-k sums S_j over a warp goes through shared memory as a
// transpose (n stores, a few 128-bit loads and a short add tree per lane) instead of 5*n shuffles.
// ---------------------------------------------------------------------------------------------
#define BMFR_FIT_WARPS (BMFR_FIT_THREADS / 32)
#define BMFR_RED_STRIDE 36  // floats per row of the transpose buffer: 16-byte aligned, bank-shifted

struct FitShared {
    float red[BMFR_FIT_WARPS][BMFR_BUFFER_COUNT][BMFR_RED_STRIDE];  // per-warp transpose buffer
    float coef[BMFR_FIT_WARPS][16];                                   // per-warp S_j / S_k broadcast
    float rstack[BMFR_FIT_WARPS * BMFR_FEATURES + 16][BMFR_BUFFER_COUNT];  // level-1 R factors, stacked (zero padded to 96)
    float rfinal[BMFR_FEATURES][BMFR_BUFFER_COUNT];                   // level-2 R
    float minmax[BMFR_FIT_WARPS][2 * BMFR_FEATURES_SCALED];
};

// One reflector (column K) of a matrix whose rows are spread over a warp, NS rows per lane.
// rrow receives row K of R (entries K..12).
template <int NS, int K>
__device__ __forceinline__ void reflector_step(float (&a)[NS][BMFR_BUFFER_COUNT], float* __restrict__ red,
                                               float* __restrict__ coefbuf, float* __restrict__ rrow, int lane) {
    constexpr int N = BMFR_BUFFER_COUNT - K;   // columns K..12 take part
    constexpr int PARTS = (N > 8) ? 2 : 4;     // lanes that share one column's 32 partial sums
    constexpr int VALS = 32 / PARTS;           // columns handled per pass: 16 or 8
    constexpr int LEN = 32 / PARTS;            // partial sums per lane: 16 or 8
    // partial S_j = sum over this lane's rows of a_k * a_j
#pragma unroll
    for (int j = 0; j < N; ++j) {
        float acc = a[0][K] * a[0][K + j];
#pragma unroll
        for (int s = 1; s < NS; ++s) acc = fmaf(a[s][K], a[s][K + j], acc);
        red[j * BMFR_RED_STRIDE + lane] = acc;
    }
    __syncwarp();
    const int jj = lane % VALS, q = lane / VALS;
    float t = 0.f;
    if (jj < N) {
        const float4* src = reinterpret_cast<const float4*>(red + jj * BMFR_RED_STRIDE + q * LEN);
        float4 v0 = src[0], v1 = src[1];
        float t0 = (v0.x + v0.y) + (v0.z + v0.w), t1 = (v1.x + v1.y) + (v1.z + v1.w);
        if (LEN == 16) {
            float4 v2 = src[2], v3 = src[3];
            t0 += (v2.x + v2.y) + (v2.z + v2.w);
            t1 += (v3.x + v3.y) + (v3.z + v3.w);
        }
        t = t0 + t1;
    }
#pragma unroll
    for (int m = VALS; m < 32; m <<= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
    const float sk = __shfl_sync(0xffffffffu, t, 0);  // S_k = |a_k|^2
    const float c = t * (1.0f / sk);                  // 2 * dot / u_length_squared of bmfr.cl:650
    if (lane < N) {
        rrow[K + lane] = c * sqrtf(sk);               // R_kj ; R_kk = sqrt(S_k) = vec_length, bmfr.cl:583
        coefbuf[lane] = c;
    }
    __syncwarp();
    float cj[16];
#pragma unroll
    for (int i = 0; i < (N + 3) / 4; ++i) {
        const float4 v = reinterpret_cast<const float4*>(coefbuf)[i];
        cj[4 * i] = v.x; cj[4 * i + 1] = v.y; cj[4 * i + 2] = v.z; cj[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int j = 1; j < N; ++j) {
#pragma unroll
        for (int s = 0; s < NS; ++s) a[s][K + j] = fmaf(-a[s][K], cj[j], a[s][K + j]);
    }
}

template <int NS, int K>
struct ReflectorLoop {
    static __device__ __forceinline__ void run(float (&a)[NS][BMFR_BUFFER_COUNT], float* red, float* coefbuf,
                                               float (*rrows)[BMFR_BUFFER_COUNT], int lane) {
        reflector_step<NS, K>(a, red, coefbuf, rrows[K], lane);
        ReflectorLoop<NS, K + 1>::run(a, red, coefbuf, rrows, lane);
    }
};
template <int NS>
struct ReflectorLoop<NS, BMFR_FEATURES> {
    static __device__ __forceinline__ void run(float (&)[NS][BMFR_BUFFER_COUNT], float*, float*,
                                               float (*)[BMFR_BUFFER_COUNT], int) {}
};

// a[s][c]: the 13 K1 values of rows tid + 256*s.  On return weights / mins_maxs / mins_inv of `group`
// are written.
__device__ __forceinline__ void block_fit(float (&a)[BMFR_ROWS_PER_THREAD][BMFR_BUFFER_COUNT], FitShared& sh,
                                          const double* __restrict__ noise, float* __restrict__ weights,
                                          float* __restrict__ mins_maxs, float* __restrict__ mins_inv, int group) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NSC = BMFR_FEATURES_SCALED, NNS = BMFR_FEATURES_NOT_SCALED;

    // (i) block min / max of the six scaled features, bmfr.cl:511-535.  min/max are exact, so the
    // tree shape is irrelevant and mins_maxs matches the reference bit for bit.
    float mn[NSC], mx[NSC];
#pragma unroll
    for (int f = 0; f < NSC; ++f) {
        mn[f] = fminf(fminf(a[0][NNS + f], a[1][NNS + f]), fminf(a[2][NNS + f], a[3][NNS + f]));
        mx[f] = fmaxf(fmaxf(a[0][NNS + f], a[1][NNS + f]), fmaxf(a[2][NNS + f], a[3][NNS + f]));
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
        for (int f = 0; f < NSC; ++f) {
            mn[f] = fminf(mn[f], __shfl_xor_sync(0xffffffffu, mn[f], m));
            mx[f] = fmaxf(mx[f], __shfl_xor_sync(0xffffffffu, mx[f], m));
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int f = 0; f < NSC; ++f) {
            sh.minmax[warp][2 * f] = mn[f];
            sh.minmax[warp][2 * f + 1] = mx[f];
        }
    }
    if (tid < 16 * BMFR_BUFFER_COUNT) (&sh.rstack[BMFR_FIT_WARPS * BMFR_FEATURES][0])[tid] = 0.f;  // padding rows
    __syncthreads();
    float inv[NSC];
#pragma unroll
    for (int f = 0; f < NSC; ++f) {
        float lo = sh.minmax[0][2 * f], hi = sh.minmax[0][2 * f + 1];
#pragma unroll
        for (int w = 1; w < BMFR_FIT_WARPS; ++w) {
            lo = fminf(lo, sh.minmax[w][2 * f]);
            hi = fmaxf(hi, sh.minmax[w][2 * f + 1]);
        }
        mn[f] = lo;
        mx[f] = hi;
        inv[f] = scale_factor(lo, hi);
    }
    if (tid < 2 * NSC) {
        mins_maxs[(size_t)group * 2 * NSC + tid] = (tid & 1) ? mx[tid >> 1] : mn[tid >> 1];
        mins_inv[(size_t)group * 2 * NSC + tid] = (tid & 1) ? inv[tid >> 1] : mn[tid >> 1];
    }

    // scale (bmfr.cl:538-541), then the noise of the first touch (bmfr.cl:623-627) on columns 1..9
    // — in fp64 like the reference's double literal NOISE_AMOUNT, rounded to fp32 once.
#pragma unroll
    for (int s = 0; s < BMFR_ROWS_PER_THREAD; ++s) {
#pragma unroll
        for (int f = 0; f < NSC; ++f) a[s][NNS + f] = scale_feature(a[s][NNS + f], mn[f], inv[f]);
#pragma unroll
        for (int c = 1; c < BMFR_FEATURES; ++c)
            a[s][c] = (float)((double)a[s][c] + __ldg(&noise[(c - 1) * BMFR_BLOCK_PIXELS + tid + BMFR_FIT_THREADS * s]));
    }

    // (ii) level 1 of the TSQR: every warp factors its own 128 rows, no block barrier
    ReflectorLoop<BMFR_ROWS_PER_THREAD, 0>::run(a, &sh.red[warp][0][0], sh.coef[warp],
                                                &sh.rstack[warp * BMFR_FEATURES], lane);
    __syncthreads();

    // level 2 + (iii) back-substitution in warp 0: 80 stacked rows (+16 zero rows), three per lane
    if (warp == 0) {
        constexpr int NS2 = 3;
        float b[NS2][BMFR_BUFFER_COUNT];
#pragma unroll
        for (int s = 0; s < NS2; ++s) {
            const int row = lane + 32 * s;
            const int k = row % BMFR_FEATURES;  // rows of a level-1 R are zero left of their diagonal
#pragma unroll
            for (int c = 0; c < BMFR_BUFFER_COUNT; ++c)
                b[s][c] = (row >= BMFR_FIT_WARPS * BMFR_FEATURES || c >= k) ? sh.rstack[row][c] : 0.f;
        }
        ReflectorLoop<NS2, 0>::run(b, &sh.red[0][0][0], sh.coef[0], sh.rfinal, lane);
        __syncwarp();
        // lane i < 10 takes row i of R; solve R x = rhs for the three channels (bmfr.cl:659-692)
        const int r = lane < BMFR_FEATURES ? lane : 0;
        float row[BMFR_BUFFER_COUNT];
#pragma unroll
        for (int c = 0; c < BMFR_BUFFER_COUNT; ++c) row[c] = sh.rfinal[r][c];
        float rhs[3] = {row[10], row[11], row[12]};
        float x[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int i = BMFR_FEATURES - 1; i >= 0; --i) {
            const float d = __shfl_sync(0xffffffffu, row[i], i);  // R_ii
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float xi = __shfl_sync(0xffffffffu, rhs[c], i) / d;
                if (lane == i) x[c] = xi;
                if (lane < i) rhs[c] = fmaf(-row[i], xi, rhs[c]);
            }
        }
        if (lane < BMFR_FEATURES) {
            float* wout = weights + ((size_t)group * BMFR_FEATURES + lane) * 3;  // bmfr.cl:694-699
            wout[0] = x[0];
            wout[1] = x[1];
            wout[2] = x[2];
        }
    }
}
