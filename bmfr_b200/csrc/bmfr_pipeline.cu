// Host pipeline + C ABI (include/bmfr_b200.h).  Replaces the OpenCL setup and the frame loop of
// /root/reference/opencl/bmfr.cpp: buffer creation (:315-343), static argument binding (:349-384),
// per-frame argument binding and the five launches (:429-476), the double-buffer swap (:483-484)
// and the per-kernel event timers (:386-397,488-506).  The arithmetic lives in bmfr_kernels.cu.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/bmfr_b200.h"
#include "bmfr_error.h"
#include "bmfr_kernels.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_last_error[512] = "";

int bmfr_set_error(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return status;
}

int bmfr_check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return BMFR_OK;
    int st = BMFR_ERR_CUDA;
    if (e == cudaErrorMemoryAllocation) st = BMFR_ERR_OUT_OF_MEMORY;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) st = BMFR_ERR_NO_DEVICE;
    return bmfr_set_error(st, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kProfileSlots = 256;  // stage timers are kept for the last kProfileSlots frames
constexpr int kHostSlots = 3;       // upload ring of the host-pointer entry: current, previous, next

// Double_buffer<T> of bmfr.cpp:122-135: two equal buffers and a swap().
template <class T>
struct DoubleBuffer {
    T* buf[2] = {nullptr, nullptr};
    bool swapped = false;
    T* current() const { return swapped ? buf[0] : buf[1]; }
    T* previous() const { return swapped ? buf[1] : buf[0]; }
    void swap() { swapped = !swapped; }
};

struct StageEvents {
    cudaEvent_t ev[7] = {};  // boundaries: before K1, after K1, K2, K3, K4, K5 (FUSED: before, after reproject, fit, post)
    int frame = -1;
    bool created = false;
};

}  // namespace

#ifndef BMFR_HISTORY_PADDED
#define BMFR_HISTORY_PADDED 0  // measured at 1080p: post 60.2 vs 55.8 us (profiles/r02_u_*): the 16 B per pixel of extra DRAM traffic cost more than the L1 wavefronts save
#endif

struct bmfr_ctx {
    bmfr_params prm;
    bmfr_geometry geo;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;

    // the reference's 15 buffers (bmfr.cpp:315-343); normals/positions/noisy inputs belong to the caller
    DoubleBuffer<float> noisy_acc;          // noisy_buffer after K1's in-place store
    DoubleBuffer<unsigned char> spp;        // spp_buffer
    DoubleBuffer<float> accum;              // out_buffer
    DoubleBuffer<float> result;             // result_buffer
    float2* prev_pixels = nullptr;          // prev_pixels_buffer
    unsigned char* accept = nullptr;        // accept_buffer
    float* tmp_data = nullptr;              // in_buffer (STAGED)
    float* filtered = nullptr;              // filtered_buffer (STAGED)
    float* tone_mapped = nullptr;           // tone_mapped_buffer (STAGED)
    float* weights = nullptr;               // weights_buffer
    float* mins_maxs = nullptr;             // mins_maxs_buffer
    float* mins_inv = nullptr;              // (min, 1/range) twin of mins_maxs used by the weighted sum
    double* noise = nullptr;
    float* noise_f = nullptr;               // the tile rounded to fp32 (FUSED fit)
    int* d_oob = nullptr;  // [0] out-of-strip flag, [1] fit block draw
    float* tri = nullptr;  // FUSED: level-1 triangles of the fit
    int tmp_block_rows = 0;

    const float* prev_normals = nullptr;    // normals_buffer.previous(): the caller's pointer of the last call
    const float* prev_positions = nullptr;
    bool has_prev = false;
    long long launches = 0;

    // host-pointer entry
    float* up[4][kHostSlots] = {};          // albedo, normal, position, noisy upload ring
    // Layout of accum / result: 3 floats per pixel (bmfr.cl:224-241), or 4 — a pixel padded to 16 bytes, so that a tap of the
    // post pass's gathers is one 128-bit load (whole-image FUSED contexts whose post pass is the TMA kernel; bmfr_post.cu).
    // Callers never see the padding: bmfr_get_buffer() hands out a compacted copy, the host entry reads the result through
    // the post pass's user_out copy.
    int hist_stride = 3;
    float* hist_compact[2] = {nullptr, nullptr};  // bmfr_get_buffer(ACCUM / RESULT) of a padded context
    float* out_stage[2] = {nullptr, nullptr};     // host entry of a padded context: the frame's result, 3 floats per pixel
    cudaEvent_t up_done[kHostSlots] = {};   // H2D of the slot finished
    cudaEvent_t frame_done[kHostSlots] = {};  // kernels that read the slot finished
    cudaEvent_t d2h_done[2] = {};
    long long host_frames = 0;
    bool host_ready = false;

    std::vector<StageEvents> prof;
    unsigned long long* d_stamps = nullptr;  // profile = 2: [kProfileSlots][3 kernels][first start, ~last end] (globaltimer ns)
    int stamp_frame[kProfileSlots];

    // peer-to-peer halo exchange (sharded contexts): the neighbour above (side 0) / below (side 1)
    struct Peer {
        bool connected = false, ipc = false, same_device = false;
        int row0 = 0, row1 = 0, own_y0 = 0, own_y1 = 0;    // the neighbour's geometry
        float* noisy_acc[2] = {nullptr, nullptr};          // its state buffers (peer-mapped), by physical index
        unsigned char* spp[2] = {nullptr, nullptr};
        float* accum[2] = {nullptr, nullptr};
        float* result[2] = {nullptr, nullptr};
        unsigned int* flags = nullptr;                     // its flag pair; this context signals flags[1 - side]
    } peer[2];
    // Written by the neighbours' GPUs: [0], [1] = frames whose accumulated noisy colour + spp rows have arrived from the
    // neighbour above / below (raised by its reprojection), [4], [5] = frames whose accumulated filtered colour + TAA
    // rows have arrived (raised by its post pass).  Local: [2] a wait timed out (sticky), [8], [9] zone-CTA counters of
    // the reprojection / the post pass.
    unsigned int* d_flags = nullptr;
    bool failed = false;               // a halo wait timed out: the temporal state is no longer trustworthy
    long long seq = 0;                 // frames submitted on this context

    // Overlapped frames (params.overlap_frames, FUSED contexts, strips included): reprojection, fit and post pass
    // each have their own stream; events order them within a frame (R -> F -> P) and across frames (R(f)
    // after P(f-2): everything a frame hands from one kernel to the next exists twice, indexed by frame
    // parity, so frame f+1 can start while frame f is still in its fit / post pass).
    struct Overlap {
        bool on = false;
        cudaStream_t s_r = nullptr, s_f = nullptr, s_p = nullptr;
        cudaEvent_t e_in[2] = {}, e_r[2] = {}, e_f[2] = {}, e_p[2] = {};
        // the second copy (odd frames) of the per-frame temporaries; even frames use the context's own
        float2* prev_pixels = nullptr;
        unsigned char* accept = nullptr;
        float *weights = nullptr, *mins_maxs = nullptr, *mins_inv = nullptr, *noise_f = nullptr;
        double* noise = nullptr;
        int* counter = nullptr;
    } ov;
    int parity() const { return ov.on ? (int)(seq & 1) : 0; }            // of the frame being submitted
    int last_parity() const { return ov.on ? (int)((seq + 1) & 1) : 0; }  // of the last submitted frame
};

static size_t rows_of(const bmfr_ctx* c) { return (size_t)(c->geo.row1 - c->geo.row0); }

template <class T>
static int dev_alloc(T** p, size_t count, const char* what) {
    cudaError_t e = cudaMalloc((void**)p, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
    if (e != cudaSuccess) return bmfr_check_cuda(e, what);
    return BMFR_OK;
}

static void free_ctx(bmfr_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->prm.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->h2d_stream) { cudaStreamSynchronize(c->h2d_stream); cudaStreamDestroy(c->h2d_stream); }
    if (c->d2h_stream) { cudaStreamSynchronize(c->d2h_stream); cudaStreamDestroy(c->d2h_stream); }
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->noisy_acc.buf[i]);
        cudaFree(c->spp.buf[i]);
        cudaFree(c->accum.buf[i]);
        cudaFree(c->result.buf[i]);
        cudaFree(c->hist_compact[i]);
        cudaFree(c->out_stage[i]);
        if (c->d2h_done[i]) cudaEventDestroy(c->d2h_done[i]);
    }
    cudaFree(c->prev_pixels);
    cudaFree(c->accept);
    cudaFree(c->tmp_data);
    cudaFree(c->filtered);
    cudaFree(c->tone_mapped);
    cudaFree(c->weights);
    cudaFree(c->mins_maxs);
    cudaFree(c->mins_inv);
    cudaFree(c->noise);
    cudaFree(c->noise_f);
    cudaFree(c->d_oob);
    cudaFree(c->tri);
    {
        bmfr_ctx::Overlap& o = c->ov;
        for (cudaStream_t st : {o.s_r, o.s_f, o.s_p})
            if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
        for (int i = 0; i < 2; ++i)
            for (cudaEvent_t e : {o.e_in[i], o.e_r[i], o.e_f[i], o.e_p[i]})
                if (e) cudaEventDestroy(e);
        cudaFree(o.prev_pixels); cudaFree(o.accept); cudaFree(o.weights); cudaFree(o.mins_maxs); cudaFree(o.mins_inv);
        cudaFree(o.noise); cudaFree(o.noise_f); cudaFree(o.counter);
    }
    for (int side = 0; side < 2; ++side) {
        bmfr_ctx::Peer& pr = c->peer[side];
        if (pr.connected && pr.ipc) {
            for (int i = 0; i < 2; ++i) {
                cudaIpcCloseMemHandle(pr.noisy_acc[i]); cudaIpcCloseMemHandle(pr.spp[i]);
                cudaIpcCloseMemHandle(pr.accum[i]); cudaIpcCloseMemHandle(pr.result[i]);
            }
            cudaIpcCloseMemHandle(pr.flags);
        }
    }
    cudaFree(c->d_flags);
    cudaFree(c->d_stamps);
    for (int k = 0; k < 4; ++k)
        for (int s = 0; s < kHostSlots; ++s) cudaFree(c->up[k][s]);
    for (int s = 0; s < kHostSlots; ++s) {
        if (c->up_done[s]) cudaEventDestroy(c->up_done[s]);
        if (c->frame_done[s]) cudaEventDestroy(c->frame_done[s]);
    }
    for (auto& p : c->prof)
        if (p.created)
            for (auto& e : p.ev) cudaEventDestroy(e);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" {

const char* bmfr_last_error(void) { return g_last_error; }
int bmfr_abi_version(void) { return BMFR_B200_ABI_VERSION; }

void bmfr_default_params(bmfr_params* p, int width, int height) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->width = width;
    p->height = height;
    p->device = 0;
    p->mode = BMFR_MODE_FUSED;
    p->noise_amount = 1e-2;          // bmfr.cpp:58
    p->blend_alpha = 0.2f;           // bmfr.cpp:60
    p->second_blend_alpha = 0.1f;    // bmfr.cpp:61
    p->taa_blend_alpha = 0.2f;       // bmfr.cpp:62
    p->position_limit_squared = 0.f;
    p->normal_limit_squared = 0.f;
    bmfr_synth_limits(&p->position_limit_squared, &p->normal_limit_squared);
    p->tmp_half = 0;
    p->profile = 0;
}

int bmfr_feature_counts(int feature_set, int* features, int* scaled) {
    static const int kF[BMFR_FEATURE_SET_COUNT_] = {10, 7, 7}, kS[BMFR_FEATURE_SET_COUNT_] = {6, 3, 6};
    if (feature_set < 0 || feature_set >= BMFR_FEATURE_SET_COUNT_)
        return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_feature_counts: unknown feature set %d", feature_set);
    if (features) *features = kF[feature_set];
    if (scaled) *scaled = kS[feature_set];
    return BMFR_OK;
}

void bmfr_block_offset(int frame, int* off_x, int* off_y) {
    int x, y;
    bmfr_host_block_offset(frame, &x, &y);
    if (off_x) *off_x = x;
    if (off_y) *off_y = y;
}

int bmfr_create(const bmfr_params* params, bmfr_ctx** out_ctx) {
    if (!params || !out_ctx) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_create: null argument");
    *out_ctx = nullptr;
    const bmfr_params& p = *params;
    // "x and y are always less than one size out of bounds if image dimensions are bigger than
    // BLOCK_EDGE_LENGTH" (bmfr.cl:312-313): mirror() needs at least one block edge of image.
    if (p.width < BMFR_BLOCK_EDGE || p.height < BMFR_BLOCK_EDGE)
        return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_create: image %dx%d smaller than one 32x32 block",
                              p.width, p.height);
    // mirror() "works only if index is less than one size out of bounds" (bmfr.cl:207-208): the margin grid
    // reaches WORKSET + 29 pixels, which must still mirror into the image.  The reference reads out of
    // bounds for such sizes; this library refuses them.
    {
        const int ww = BMFR_BLOCK_EDGE * ((p.width + BMFR_BLOCK_EDGE - 1) / BMFR_BLOCK_EDGE);
        const int hw = BMFR_BLOCK_EDGE * ((p.height + BMFR_BLOCK_EDGE - 1) / BMFR_BLOCK_EDGE);
        if (ww + 29 > 2 * p.width - 1 || hw + 29 > 2 * p.height - 1)
            return bmfr_set_error(BMFR_ERR_UNSUPPORTED,
                                  "bmfr_create: image %dx%d: the 32-pixel margin of the block grid cannot be mirrored into the "
                                  "image (needs 2*size - 1 >= workset + 29, bmfr.cl:207-216)", p.width, p.height);
    }
    if (p.mode != BMFR_MODE_STAGED && p.mode != BMFR_MODE_FUSED)
        return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_create: unknown mode %d", p.mode);
    if (p.fit_method != BMFR_FIT_GRAM && p.fit_method != BMFR_FIT_TSQR)
        return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_create: unknown fit_method %d", p.fit_method);
    if (p.halo_timeout_ms < 0) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_create: negative halo_timeout_ms");
    if (p.feature_set < 0 || p.feature_set >= BMFR_FEATURE_SET_COUNT_)
        return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_create: unknown feature_set %d", p.feature_set);
    if (p.feature_set != BMFR_FEATURE_SET_DEFAULT && (p.mode != BMFR_MODE_FUSED || p.fit_method != BMFR_FIT_GRAM))
        return bmfr_set_error(BMFR_ERR_UNSUPPORTED, "bmfr_create: feature lists other than the reference's shipped one are instantiated "
                                                    "for the FUSED kernels with fit_method = BMFR_FIT_GRAM only");
    if ((p.tmp_half != 0 || p.reference_order != 0) && p.mode != BMFR_MODE_STAGED)
        return bmfr_set_error(BMFR_ERR_UNSUPPORTED,
                              "bmfr_create: tmp_half / reference_order are the STAGED compatibility path (the FUSED fit is fp32 and "
                              "keeps tmp_data in registers)");
    const bool whole = (p.strip_y0 == 0 && p.strip_y1 == 0);
    if (!whole && (p.strip_y0 < 0 || p.strip_y1 > p.height || p.strip_y0 >= p.strip_y1 || p.halo_rows < 0))
        return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_create: bad strip [%d,%d) halo %d", p.strip_y0,
                              p.strip_y1, p.halo_rows);

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return bmfr_set_error(BMFR_ERR_NO_DEVICE, "bmfr_create: no CUDA device (%s); this library has no CPU path",
                              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (p.device < 0 || p.device >= ndev)
        return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_create: device %d of %d", p.device, ndev);
    BMFR_CUDA_TRY(cudaSetDevice(p.device));

    bmfr_ctx* c = new (std::nothrow) bmfr_ctx();
    if (!c) return bmfr_set_error(BMFR_ERR_OUT_OF_MEMORY, "bmfr_create: host allocation failed");
    c->prm = p;
    if (c->prm.tmp_half) c->prm.reference_order = 1;  // fp16 rounding of every tmp_data store only exists in the reference's schedule
    bmfr_geometry& g = c->geo;
    g.width = p.width;
    g.height = p.height;
    g.workset_width = BMFR_BLOCK_EDGE * ((p.width + BMFR_BLOCK_EDGE - 1) / BMFR_BLOCK_EDGE);    // bmfr.cpp:107-108
    g.workset_height = BMFR_BLOCK_EDGE * ((p.height + BMFR_BLOCK_EDGE - 1) / BMFR_BLOCK_EDGE);  // bmfr.cpp:109-110
    g.margin_width = g.workset_width + BMFR_BLOCK_EDGE;                                         // bmfr.cpp:111
    g.margin_height = g.workset_height + BMFR_BLOCK_EDGE;                                       // bmfr.cpp:112
    g.blocks_x = g.margin_width / BMFR_BLOCK_EDGE;
    g.blocks_y = g.margin_height / BMFR_BLOCK_EDGE;
    if (whole) {
        g.own_y0 = 0; g.own_y1 = p.height; g.row0 = 0; g.row1 = p.height;
        g.block_row0 = 0; g.block_row1 = g.blocks_y;
    } else {
        g.own_y0 = p.strip_y0; g.own_y1 = p.strip_y1;
        g.row0 = p.strip_y0 - p.halo_rows < 0 ? 0 : p.strip_y0 - p.halo_rows;
        g.row1 = p.strip_y1 + p.halo_rows > p.height ? p.height : p.strip_y1 + p.halo_rows;
        const int py0 = g.own_y0 > 0 ? g.own_y0 - 1 : 0, py1 = g.own_y1 < p.height ? g.own_y1 + 1 : p.height;
        g.block_row0 = (py0 + 16 - 14) >> 5;            // largest offset.y is 14, smallest -16 (bmfr.cl:268-285)
        g.block_row1 = ((py1 - 1 + 16 + 16) >> 5) + 1;
        if (g.block_row1 > g.blocks_y) g.block_row1 = g.blocks_y;
    }
    c->tmp_block_rows = g.block_row1 - g.block_row0;

    int st = BMFR_OK;
    auto fail = [&](int s) { free_ctx(c); return s; };
    if (p.stream) {
        c->stream = (cudaStream_t)p.stream;
    } else {
        if ((st = bmfr_check_cuda(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "stream")) != 0) return fail(st);
        c->own_stream = true;
    }
    const size_t npix = rows_of(c) * (size_t)p.width;
    const size_t nb = (size_t)g.blocks_x * g.blocks_y;
    const bool whole_image = g.row0 == 0 && g.row1 == p.height;
    c->hist_stride = (BMFR_HISTORY_PADDED && p.mode == BMFR_MODE_FUSED && whole_image && post_uses_tma(p.width, p.height)) ? 4 : 3;
    for (int i = 0; i < 2 && st == 0; ++i) {
        if (st == 0) st = dev_alloc(&c->noisy_acc.buf[i], npix * 3, "noisy_acc");
        if (st == 0) st = dev_alloc(&c->spp.buf[i], npix, "spp");
        if (st == 0) st = dev_alloc(&c->accum.buf[i], npix * c->hist_stride, "accum");
        if (st == 0) st = dev_alloc(&c->result.buf[i], npix * c->hist_stride, "result");
    }
    if (st == 0) st = dev_alloc(&c->prev_pixels, npix, "prev_pixels");
    if (st == 0) st = dev_alloc(&c->accept, npix, "accept");
    if (st == 0) st = dev_alloc(&c->weights, nb * BMFR_FEATURES * 3, "weights");
    if (st == 0) st = dev_alloc(&c->mins_maxs, nb * BMFR_FEATURES_SCALED * 2, "mins_maxs");
    if (st == 0) st = dev_alloc(&c->mins_inv, nb * BMFR_FEATURES_SCALED * 2, "mins_inv");
    if (st == 0) st = dev_alloc(&c->noise, (size_t)(BMFR_FEATURES - 1) * BMFR_BLOCK_PIXELS, "noise");
    if (st == 0) st = dev_alloc(&c->noise_f, (size_t)(BMFR_FEATURES - 1) * BMFR_BLOCK_PIXELS, "noise_f");
    if (st == 0) st = dev_alloc(&c->d_oob, 2, "oob flag + block counter");
    if (st == 0 && p.mode == BMFR_MODE_FUSED) st = dev_alloc(&c->tri, nb * 4 * 136, "level-1 triangles");
    if (st == 0) st = dev_alloc(&c->d_flags, 16, "halo flags");
    if (st == 0 && p.mode == BMFR_MODE_STAGED) {
        st = dev_alloc(&c->tmp_data, (size_t)c->tmp_block_rows * g.blocks_x * BMFR_BUFFER_COUNT * BMFR_BLOCK_PIXELS, "tmp_data");
        if (st == 0) st = dev_alloc(&c->filtered, npix * 3, "filtered");
        if (st == 0) st = dev_alloc(&c->tone_mapped, npix * 3, "tone_mapped");
    }
    if (st != 0) return fail(st);
    if ((st = bmfr_check_cuda(cudaMemsetAsync(c->d_oob, 0, 2 * sizeof(int), c->stream), "memset")) != 0) return fail(st);
    if ((st = bmfr_check_cuda(cudaMemsetAsync(c->d_flags, 0, 16 * sizeof(unsigned int), c->stream), "memset")) != 0) return fail(st);
    // weights of blocks a strip never fits stay defined
    if ((st = bmfr_check_cuda(cudaMemsetAsync(c->weights, 0, nb * BMFR_FEATURES * 3 * sizeof(float), c->stream), "memset")) != 0) return fail(st);
    if ((st = bmfr_check_cuda(cudaMemsetAsync(c->mins_maxs, 0, nb * BMFR_FEATURES_SCALED * 2 * sizeof(float), c->stream), "memset")) != 0) return fail(st);
    if ((st = bmfr_check_cuda(cudaMemsetAsync(c->mins_inv, 0, nb * BMFR_FEATURES_SCALED * 2 * sizeof(float), c->stream), "memset")) != 0) return fail(st);
    if (p.profile == 1) c->prof.resize(kProfileSlots);
    if (p.profile == 2 && p.mode == BMFR_MODE_FUSED) {
        if ((st = dev_alloc(&c->d_stamps, (size_t)kProfileSlots * 6, "kernel stamps")) != 0) return fail(st);
        if ((st = bmfr_check_cuda(cudaMemsetAsync(c->d_stamps, 0xFF, (size_t)kProfileSlots * 6 * sizeof(unsigned long long), c->stream), "memset")) != 0) return fail(st);
        for (int i = 0; i < kProfileSlots; ++i) c->stamp_frame[i] = -1;
    }
    if (p.overlap_frames && p.mode == BMFR_MODE_FUSED && p.profile != 1) {
        bmfr_ctx::Overlap& o = c->ov;
        if (st == 0) st = dev_alloc(&o.prev_pixels, npix, "prev_pixels (odd frames)");
        if (st == 0) st = dev_alloc(&o.accept, npix, "accept (odd frames)");
        if (st == 0) st = dev_alloc(&o.weights, nb * BMFR_FEATURES * 3, "weights (odd frames)");
        if (st == 0) st = dev_alloc(&o.mins_maxs, nb * BMFR_FEATURES_SCALED * 2, "mins_maxs (odd frames)");
        if (st == 0) st = dev_alloc(&o.mins_inv, nb * BMFR_FEATURES_SCALED * 2, "mins_inv (odd frames)");
        if (st == 0) st = dev_alloc(&o.noise, (size_t)(BMFR_FEATURES - 1) * BMFR_BLOCK_PIXELS, "noise (odd frames)");
        if (st == 0) st = dev_alloc(&o.noise_f, (size_t)(BMFR_FEATURES - 1) * BMFR_BLOCK_PIXELS, "noise_f (odd frames)");
        if (st == 0) st = dev_alloc(&o.counter, 1, "block counter (odd frames)");
        // (stream priorities — the post pass over the fit over the reprojection, or the other way round — both measured 8.5 k
        // against 9.3 k frames/s at 1080p: the hardware's own interleaving of the three grids is the better schedule)
        for (cudaStream_t* ps : {&o.s_r, &o.s_f, &o.s_p})
            if (st == 0) st = bmfr_check_cuda(cudaStreamCreateWithFlags(ps, cudaStreamNonBlocking), "cudaStreamCreate");
        for (int i = 0; i < 2; ++i)
            for (cudaEvent_t* pe : {&o.e_in[i], &o.e_r[i], &o.e_f[i], &o.e_p[i]})
                if (st == 0) st = bmfr_check_cuda(cudaEventCreateWithFlags(pe, cudaEventDisableTiming), "cudaEventCreate");
        if (st != 0) return fail(st);
        o.on = true;
    }
    if ((st = bmfr_check_cuda(cudaStreamSynchronize(c->stream), "create sync")) != 0) return fail(st);
    *out_ctx = c;
    return BMFR_OK;
}

void bmfr_destroy(bmfr_ctx* ctx) { free_ctx(ctx); }

int bmfr_get_geometry(const bmfr_ctx* ctx, bmfr_geometry* out) {
    if (!ctx || !out) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_geometry: null argument");
    *out = ctx->geo;
    return BMFR_OK;
}

long long bmfr_kernel_launches(const bmfr_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

static bool halo_active(const bmfr_ctx* c) { return c->peer[0].connected || c->peer[1].connected; }
static bool halo_rows_for(const bmfr_ctx* c, int side, bool late, int* y0, int* y1);

// The halo-exchange duties of this frame's reprojection (early: accumulated noisy colour + spp) and post pass (late:
// accumulated filtered colour + TAA result); see HaloK in bmfr_kernels.h and DESIGN.md 5 for the protocol:
//   R(f) waits for early >= f (the rows it gathers from have arrived; and the neighbours' R(f-1) zone no longer reads the
//        halo rows this frame's pushes overwrite — same physical buffer, two frames apart) and late >= f-1 (their fit /
//        post pass of frame f-2, which read that buffer as "current", are done); raises early to f+1;
//   P(f) waits for late >= f (history rows arrived; their P(f-1) zone no longer reads what this frame's pushes
//        overwrite); raises late to f+1.
// Every wait refers to frames submitted earlier on both sides, so the order of submission is a schedule that never
// blocks — in order (one stream) as well as with overlapped frames (three streams).
static void fill_halo(bmfr_ctx* c, KParams& P) {
    if (!halo_active(c) || c->prm.mode != BMFR_MODE_FUSED) return;
    const bmfr_geometry& g = c->geo;
    const unsigned int f = (unsigned int)c->seq;
    const int cur = c->noisy_acc.swapped ? 0 : 1;  // physical index of the buffers this frame writes (all four swap together)
    const unsigned long long timeout_ns = (unsigned long long)(c->prm.halo_timeout_ms > 0 ? c->prm.halo_timeout_ms : 10000) * 1000000ull;
    for (int late = 0; late < 2; ++late) {
        HaloK& h = late ? P.halo_p : P.halo_r;
        h.active = 1;
        h.flags = c->d_flags;
        h.wait_early = late ? 0u : f;
        h.wait_late = late ? f : (f >= 1 ? f - 1 : 0u);
        h.signal_value = f + 1;
        h.done_counter = c->d_flags + 8 + late;
        h.timeout_ns = timeout_ns;
        // The zone: rows whose CTAs can gather from halo rows or own rows a neighbour mirrors.  The reprojection mirrors
        // halo_rows rows; the post pass gathers from, and mirrors, halo_rows - 32 (halo_rows_for) around the owned rows +- 1.
        const int h2 = c->prm.halo_rows > 32 ? c->prm.halo_rows - 32 : 0;
        const int zone_rows = late ? (h2 + 2 < c->prm.halo_rows ? h2 + 2 : c->prm.halo_rows) : c->prm.halo_rows;
        h.zone_y[0] = c->peer[0].connected ? g.own_y0 + zone_rows : -(1 << 30);
        h.zone_y[1] = c->peer[1].connected ? g.own_y1 - zone_rows : (1 << 30);
        for (int s = 0; s < 2; ++s) {
            const bmfr_ctx::Peer& pr = c->peer[s];
            h.side_on[s] = pr.connected ? 1 : 0;
            h.push_y0[s] = h.push_y1[s] = 0;
            if (!pr.connected) continue;
            int y0, y1;
            if (halo_rows_for(c, s, late != 0, &y0, &y1)) {
                if (!late) {
                    // The rows next to the edge that the neighbour reprojects itself this frame — the rows of its blocks that
                    // straddle the edge, bit-identical to ours by construction — need not travel: they are the rows of its
                    // first / last block row of this frame (KParams::k1_y0 / k1_y1 on its side, from this frame's block offset).
                    if (s == 1) {  // neighbour below: it covers [k1_y0', own_y1) itself
                        const int nb_by0 = (g.own_y1 - 1 + 16 - P.off_y) >> 5;
                        const int nb_k1_y0 = nb_by0 * 32 - 16 + P.off_y;
                        if (nb_k1_y0 < y1) y1 = nb_k1_y0 > y0 ? nb_k1_y0 : y0;
                    } else {       // neighbour above: it covers [own_y0, k1_y1') itself
                        const int nb_by1 = ((g.own_y0 + 16 - P.off_y) >> 5) + 1;
                        const int nb_k1_y1 = nb_by1 * 32 - 16 + P.off_y;
                        if (nb_k1_y1 > y0) y0 = nb_k1_y1 < y1 ? nb_k1_y1 : y1;
                    }
                }
                h.push_y0[s] = y0; h.push_y1[s] = y1;
            }
            h.peer_row0[s] = pr.row0;
            h.peer_a[s] = late ? pr.accum[cur] : pr.noisy_acc[cur];
            h.peer_b[s] = late ? pr.result[cur] : nullptr;
            h.peer_c[s] = late ? nullptr : pr.spp[cur];
            // the neighbour above sees this context as its "below" neighbour (flag 1) and vice versa
            h.peer_flag[s] = pr.flags + (late ? 4 : 0) + (s == 0 ? 1 : 0);
        }
        // zone CTAs of the launch, counted the way the kernels decide it (halo_in_zone)
        // (a CTA row inside both ends' zones — a strip shorter than two halos — counts as a top row)
        unsigned int n = 0;
        h.rows_top = h.rows_bot = 0;
        auto count = [&](int ya, int yb) {
            if (ya < h.zone_y[0]) { ++h.rows_top; ++n; }
            else if (yb > h.zone_y[1]) { ++h.rows_bot; ++n; }
        };
        if (!late) {
            for (int ya = P.k1_y0; ya < P.k1_y1; ya += 32) count(ya, ya + 32);
            n *= (unsigned int)((g.width + 31) / 32);
        } else {
            for (int by = P.by0; by < P.by1; ++by) {
                const int y0 = by * 32 - 16 + P.off_y;
                count(y0 - 1, y0 + 33);
            }
            n *= (unsigned int)P.blocks_x;
        }
        h.zone_ctas = n;
    }
}

// Per-frame argument binding, bmfr.cpp:429-474.
static void fill_params(bmfr_ctx* c, KParams& P, int frame, const float* d_albedo, const float* d_normal,
                        const float* d_position, const float* d_noisy, const float* cam_prev, const float* pixel_offset,
                        float* d_out) {
    const bmfr_geometry& g = c->geo;
    memset(&P, 0, sizeof(P));
    P.W = g.width; P.H = g.height; P.row0 = g.row0; P.row1 = g.row1; P.frame = frame;
    bmfr_host_block_offset(frame, &P.off_x, &P.off_y);
    P.blocks_x = g.blocks_x; P.blocks_y = g.blocks_y;
    P.own_y0 = g.own_y0; P.own_y1 = g.own_y1;
    const bool whole = (g.own_y0 == 0 && g.own_y1 == g.height);
    if (whole) {
        P.py0 = 0; P.py1 = g.height; P.by0 = 0; P.by1 = g.blocks_y;
    } else {
        P.py0 = g.own_y0 > 0 ? g.own_y0 - 1 : 0;
        P.py1 = g.own_y1 < g.height ? g.own_y1 + 1 : g.height;
        P.by0 = (P.py0 + 16 - P.off_y) >> 5;
        P.by1 = ((P.py1 - 1 + 16 - P.off_y) >> 5) + 1;
    }
    // rows of the accumulated filtered colour / TAA result a neighbour keeps fresh (see halo_rows_for)
    {
        const int h2 = c->prm.halo_rows > 32 ? c->prm.halo_rows - 32 : 0;
        P.state2_row0 = whole ? 0 : (g.own_y0 - h2 > g.row0 ? g.own_y0 - h2 : g.row0);
        P.state2_row1 = whole ? g.height : (g.own_y1 + h2 < g.row1 ? g.own_y1 + h2 : g.row1);
    }
    // rows the blocks by0..by1 cover (mirrored margin rows fold back into the first / last block row)
    P.k1_y0 = P.by0 * 32 - 16 + P.off_y; if (P.k1_y0 < 0) P.k1_y0 = 0;
    P.k1_y1 = P.by1 * 32 - 16 + P.off_y; if (P.k1_y1 > g.height) P.k1_y1 = g.height;
    if (cam_prev) memcpy(P.cam, cam_prev, 16 * sizeof(float));
    P.poff_x = pixel_offset ? pixel_offset[0] : 0.f;
    P.poff_y1 = 1.f - (pixel_offset ? pixel_offset[1] : 0.f);
    P.blend_alpha = c->prm.blend_alpha;
    P.second_blend_alpha = c->prm.second_blend_alpha;
    P.taa_blend_alpha = c->prm.taa_blend_alpha;
    P.pos_limit = c->prm.position_limit_squared;
    P.nrm_limit = c->prm.normal_limit_squared;
    P.cur_normals = d_normal; P.prev_normals = c->prev_normals;
    P.cur_positions = d_position; P.prev_positions = c->prev_positions;
    P.cur_noisy = d_noisy;
    P.prev_noisy_acc = c->noisy_acc.previous(); P.cur_noisy_acc = c->noisy_acc.current();
    P.prev_spp = c->spp.previous(); P.cur_spp = c->spp.current();
    const bool odd = c->parity() == 1;
    P.prev_pixels = odd ? c->ov.prev_pixels : c->prev_pixels; P.accept = odd ? c->ov.accept : c->accept;
    P.tmp_half = c->prm.tmp_half; P.reference_order = c->prm.reference_order;
    P.tmp_data = c->tmp_data;
    P.weights = odd ? c->ov.weights : c->weights;
    P.mins_maxs = odd ? c->ov.mins_maxs : c->mins_maxs;
    P.mins_inv = odd ? c->ov.mins_inv : c->mins_inv;
    P.noise = odd ? c->ov.noise : c->noise;
    P.noise_f = odd ? c->ov.noise_f : c->noise_f;
    P.noise_out = const_cast<double*>(P.noise); P.noise_f_out = const_cast<float*>(P.noise_f); P.noise_amount = c->prm.noise_amount;
    P.albedo = d_albedo; P.filtered = c->filtered;
    P.accum_prev = c->accum.previous(); P.accum_cur = c->accum.current();
    P.tone_mapped = c->tone_mapped;
    P.result_prev = c->result.previous(); P.result_cur = c->result.current();
    P.user_out = d_out; P.oob_flag = c->d_oob; P.block_counter = odd ? c->ov.counter : c->d_oob + 1;
    P.plain_launch = c->ov.on ? 1 : 0;
    P.hist_stride = c->hist_stride;
    P.fit_method = c->prm.fit_method;
    P.feature_set = c->prm.feature_set;
    bmfr_feature_counts(c->prm.feature_set, &P.n_features, &P.n_scaled);
    P.tri = c->tri;
    if (c->d_stamps) {  // this frame's slot was armed by the previous frame's reprojection (the first one at create)
        const size_t slot = (size_t)c->seq % kProfileSlots, next = (size_t)(c->seq + 1) % kProfileSlots;
        P.stamps = c->d_stamps + slot * 6;
        P.stamps_next = c->d_stamps + next * 6;
        c->stamp_frame[slot] = frame;
    }
    fill_halo(c, P);
}

static StageEvents* prof_slot(bmfr_ctx* c, int frame) {
    if (c->prof.empty()) return nullptr;
    StageEvents& s = c->prof[(size_t)(frame >= 0 ? frame : 0) % kProfileSlots];
    if (!s.created) {
        for (auto& e : s.ev)
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        s.created = true;
    }
    s.frame = frame;
    return &s;
}

#define LAUNCH_TRY(call, name)                              \
    do {                                                    \
        int _st = bmfr_check_cuda((call), name);            \
        if (_st != 0) return _st;                           \
        ++c->launches;                                      \
    } while (0)
#define MARK(i)                                             \
    do {                                                    \
        if (pe) BMFR_CUDA_TRY(cudaEventRecord(pe->ev[i], c->stream)); \
    } while (0)

static int halo_wait_same_device(bmfr_ctx* c, cudaStream_t st, unsigned int early_value, unsigned int late_value);

static int run_frame(bmfr_ctx* c, const KParams& P, int frame) {
    StageEvents* pe = prof_slot(c, frame);
    if (c->prm.mode == BMFR_MODE_STAGED)  // FUSED: the reproject kernel produces the tile itself
        LAUNCH_TRY(launch_noise_tile(c->noise, c->noise_f, c->d_oob + 1, c->prm.noise_amount, frame, c->stream), "noise_tile_kernel");
    MARK(0);
    if (c->prm.mode == BMFR_MODE_STAGED) {
        LAUNCH_TRY(launch_k1(P, c->stream), "accumulate_noisy_data");
        MARK(1);
        if (c->prm.reference_order) LAUNCH_TRY(launch_k2_reference_order(P, c->prm.tmp_half != 0, c->stream), "fitter (reference order)");
        else LAUNCH_TRY(launch_k2(P, c->stream), "fitter");
        MARK(2);
        if (c->prm.reference_order) LAUNCH_TRY(launch_k3_reference_order(P, c->stream), "weighted_sum (reference order)");
        else LAUNCH_TRY(launch_k3(P, c->stream), "weighted_sum");
        MARK(3);
        LAUNCH_TRY(launch_k4(P, c->stream), "accumulate_filtered_data");
        MARK(4);
        LAUNCH_TRY(launch_k5(P, c->stream), "taa");
        MARK(5);
    } else if (c->ov.on) {
        // R(f) | F(f) | P(f) on three streams.  R(f) follows the caller's work on the context's stream and
        // P(f-2): the buffers of this parity (accumulated colour + spp as "current", prev_pixels, accept,
        // noise tile, block counter, weights, min/max) were last read by frame f-2.  R(f-1) precedes R(f) on
        // the same stream (temporal state), as F(f-1) precedes F(f) (scratch of the fit) and P(f-1) P(f)
        // (accumulated filtered colour, TAA history).  Strips: the halo exchange happens inside R and P (fill_halo).
        bmfr_ctx::Overlap& o = c->ov;
        const int q = c->parity();
        BMFR_CUDA_TRY(cudaEventRecord(o.e_in[q], c->stream));
        BMFR_CUDA_TRY(cudaStreamWaitEvent(o.s_r, o.e_in[q], 0));
        if (c->seq >= 2) BMFR_CUDA_TRY(cudaStreamWaitEvent(o.s_r, o.e_p[q], 0));
        { int hs = halo_wait_same_device(c, o.s_r, P.halo_r.wait_early, P.halo_r.wait_late); if (hs != 0) return hs; }
        LAUNCH_TRY(launch_reproject(P, o.s_r), "reproject_kernel");
        BMFR_CUDA_TRY(cudaEventRecord(o.e_r[q], o.s_r));
        BMFR_CUDA_TRY(cudaStreamWaitEvent(o.s_f, o.e_r[q], 0));
        LAUNCH_TRY(launch_fit_qr(P, o.s_f), "fit kernel");
        BMFR_CUDA_TRY(cudaEventRecord(o.e_f[q], o.s_f));
        BMFR_CUDA_TRY(cudaStreamWaitEvent(o.s_p, o.e_f[q], 0));
        { int hs = halo_wait_same_device(c, o.s_p, 0, P.halo_p.wait_late); if (hs != 0) return hs; }
        LAUNCH_TRY(launch_post(P, o.s_p), "post_kernel");
        BMFR_CUDA_TRY(cudaEventRecord(o.e_p[q], o.s_p));
    } else {
        { int hs = halo_wait_same_device(c, c->stream, P.halo_r.wait_early, P.halo_r.wait_late); if (hs != 0) return hs; }
        LAUNCH_TRY(launch_reproject(P, c->stream), "reproject_kernel");
        MARK(1);
        LAUNCH_TRY(launch_fit_qr(P, c->stream), "fit kernel");
        MARK(2);
        { int hs = halo_wait_same_device(c, c->stream, 0, P.halo_p.wait_late); if (hs != 0) return hs; }
        LAUNCH_TRY(launch_post(P, c->stream), "post_kernel");
        MARK(3);
    }
    return BMFR_OK;
}

// ------------------------------------------------------------------------------------------------
// Peer-to-peer halo exchange (option A).  The exchange itself lives in the kernels (HaloK, fill_halo):
// zone CTAs of the reprojection / the post pass poll this context's flags in their prologue, store the rows a
// neighbour mirrors a second time into its memory (peer-mapped, NVLink) and the last of them raises the neighbour's
// flag — no copy, signal or wait launches.
//
// One exception.  Contexts connected INSIDE one process on ONE device (how the strip logic is tested on a single-GPU
// box) must not spin inside a full-size grid: the kernel they wait for may not be resident yet, and nothing guarantees
// that two grids of one device run at the same time (B200_PROFILING.md).  For such neighbours a one-thread kernel does
// the waiting in front of the launch — it occupies one CTA slot only — and the in-kernel poll then passes at once.
// ------------------------------------------------------------------------------------------------
__global__ void halo_wait_split_kernel(unsigned int* flags, int need_a, int need_b, unsigned int early_value, unsigned int late_value,
                                       unsigned long long timeout_ns) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        volatile unsigned int* f = flags;
        const bool a = !need_a || (f[0] >= early_value && f[4] >= late_value);
        const bool b = !need_b || (f[1] >= early_value && f[5] >= late_value);
        if (a && b) break;
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) {
            flags[2] = 1;  // sticky: the context has failed
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

static int halo_wait_same_device(bmfr_ctx* c, cudaStream_t st, unsigned int early_value, unsigned int late_value) {
    if (early_value == 0 && late_value == 0) return BMFR_OK;
    const int need_a = c->peer[0].connected && c->peer[0].same_device, need_b = c->peer[1].connected && c->peer[1].same_device;
    if (!need_a && !need_b) return BMFR_OK;
    const unsigned long long timeout_ns = (unsigned long long)(c->prm.halo_timeout_ms > 0 ? c->prm.halo_timeout_ms : 10000) * 1000000ull;
    halo_wait_split_kernel<<<1, 1, 0, st>>>(c->d_flags, need_a, need_b, early_value, late_value, timeout_ns);
    int rc = bmfr_check_cuda(cudaGetLastError(), "halo_wait_split_kernel");
    if (rc == 0) ++c->launches;
    return rc;
}

// Rows of a neighbour's halo this context must refresh, per state buffer.  The accumulated noisy colour and
// spp are gathered by the reprojection of every row a straddling block covers (halo_rows = 34 + motion);
// the accumulated filtered colour and the TAA result are gathered only from the owned rows +- 1, so their
// halo needs halo_rows - 32 rows (KParams::state2_row0/1 make a gather beyond that fail loudly).
static bool halo_rows_for(const bmfr_ctx* c, int side, bool late, int* y0, int* y1) {
    const bmfr_geometry& g = c->geo;
    const bmfr_ctx::Peer& pr = c->peer[side];
    int lo = pr.row0, hi = pr.row1;  // rows the neighbour stores
    if (late) {
        const int h2 = c->prm.halo_rows > 32 ? c->prm.halo_rows - 32 : 0;
        lo = pr.own_y0 - h2 > lo ? pr.own_y0 - h2 : lo;
        hi = pr.own_y1 + h2 < hi ? pr.own_y1 + h2 : hi;
    }
    if (side == 0) { *y0 = g.own_y0; *y1 = hi < g.own_y1 ? hi : g.own_y1; }
    else { *y0 = lo > g.own_y0 ? lo : g.own_y0; *y1 = g.own_y1; }
    return *y0 < *y1;
}

struct HaloBlob {  // what a neighbour needs to address this context's state: geometry + IPC handles
    unsigned int magic;
    int device, width, height, row0, row1, own_y0, own_y1, swapped;
    int mode;
    unsigned char uuid[16];  // of the device: a neighbour on the SAME device is waited for by a one-thread kernel, not in-kernel
    long long seq;
    cudaIpcMemHandle_t noisy_acc[2], spp[2], accum[2], result[2], flags;
};
static const unsigned int kHaloMagic = 0x424d4648u;  // "BMFH"

extern "C" {

int bmfr_denoise_frame(bmfr_ctx* c, int frame, const float* d_albedo, const float* d_normal, const float* d_position,
                       const float* d_noisy, const float cam_prev[16], const float pixel_offset[2], float* d_out) {
    if (!c || !d_albedo || !d_normal || !d_position || !d_noisy || !pixel_offset || frame < 0)
        return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_denoise_frame: null argument or negative frame");
    if (frame > 0 && !cam_prev)
        return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_denoise_frame: frame %d needs cam_prev", frame);
    if (frame > 0 && !c->has_prev)
        return bmfr_set_error(BMFR_ERR_SEQUENCE, "bmfr_denoise_frame: frame %d submitted before any frame 0", frame);
    BMFR_CUDA_TRY(cudaSetDevice(c->prm.device));
    if (c->failed)
        return bmfr_set_error(BMFR_ERR_SEQUENCE, "bmfr_denoise_frame: a halo wait of this context timed out earlier; its temporal state is "
                                                 "not valid any more (destroy it and start the strips again)");
    KParams P;
    fill_params(c, P, frame, d_albedo, d_normal, d_position, d_noisy, cam_prev, pixel_offset, d_out);
    int st = run_frame(c, P, frame);
    if (st != 0) return st;
    ++c->seq;
    // swap all double buffers, bmfr.cpp:483-484
    c->noisy_acc.swap(); c->spp.swap(); c->accum.swap(); c->result.swap();
    c->prev_normals = d_normal;
    c->prev_positions = d_position;
    c->has_prev = true;
    return BMFR_OK;
}

static int host_path_init(bmfr_ctx* c) {
    if (c->host_ready) return BMFR_OK;
    const size_t n = rows_of(c) * (size_t)c->geo.width * 3;
    for (int k = 0; k < 4; ++k)
        for (int s = 0; s < kHostSlots; ++s) {
            int st = dev_alloc(&c->up[k][s], n, "upload ring");
            if (st != 0) return st;
        }
    BMFR_CUDA_TRY(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
    BMFR_CUDA_TRY(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    for (int s = 0; s < kHostSlots; ++s) {
        BMFR_CUDA_TRY(cudaEventCreateWithFlags(&c->up_done[s], cudaEventDisableTiming));
        BMFR_CUDA_TRY(cudaEventCreateWithFlags(&c->frame_done[s], cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; ++i) BMFR_CUDA_TRY(cudaEventCreateWithFlags(&c->d2h_done[i], cudaEventDisableTiming));
    if (c->hist_stride != 3)
        for (int i = 0; i < 2; ++i) {
            int st = dev_alloc(&c->out_stage[i], n, "result staging");
            if (st != 0) return st;
        }
    c->host_ready = true;
    return BMFR_OK;
}

int bmfr_denoise_frame_host(bmfr_ctx* c, int frame, const float* h_albedo, const float* h_normal,
                            const float* h_position, const float* h_noisy, const float cam_prev[16],
                            const float pixel_offset[2], float* h_out) {
    if (!c || !h_albedo || !h_normal || !h_position || !h_noisy || !pixel_offset)
        return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_denoise_frame_host: null argument");
    BMFR_CUDA_TRY(cudaSetDevice(c->prm.device));
    int st = host_path_init(c);
    if (st != 0) return st;
    const size_t bytes = rows_of(c) * (size_t)c->geo.width * 3 * sizeof(float);
    const int slot = (int)(c->host_frames % kHostSlots);
    // the slot was last read by the kernels of the call three frames ago (as "previous" two ago)
    if (c->host_frames >= kHostSlots) BMFR_CUDA_TRY(cudaStreamWaitEvent(c->h2d_stream, c->frame_done[slot], 0));
    const float* src[4] = {h_albedo, h_normal, h_position, h_noisy};
    for (int k = 0; k < 4; ++k)  // bmfr.cpp:420-427
        BMFR_CUDA_TRY(cudaMemcpyAsync(c->up[k][slot], src[k], bytes, cudaMemcpyHostToDevice, c->h2d_stream));
    BMFR_CUDA_TRY(cudaEventRecord(c->up_done[slot], c->h2d_stream));
    BMFR_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->up_done[slot], 0));
    // result.current() of this frame was last read by the D2H of two frames ago
    const int rslot = (int)(c->host_frames & 1);
    // (overlapped frames: the post pass, which writes it, runs on its own stream; that stream also carries
    // "the kernels of this frame are done")
    cudaStream_t done_stream = c->ov.on ? c->ov.s_p : c->stream;
    if (c->host_frames >= 2) BMFR_CUDA_TRY(cudaStreamWaitEvent(done_stream, c->d2h_done[rslot], 0));
    // (a padded context hands the frame's result out through the post pass's user_out copy)
    float* d_result = c->hist_stride == 3 ? c->result.current() : c->out_stage[rslot];
    st = bmfr_denoise_frame(c, frame, c->up[0][slot], c->up[1][slot], c->up[2][slot], c->up[3][slot], cam_prev,
                            pixel_offset, c->hist_stride == 3 ? nullptr : d_result);
    if (st != 0) return st;
    // the previous frame's slot is released once this frame's kernels (which read it as "previous") are done
    BMFR_CUDA_TRY(cudaEventRecord(c->frame_done[slot], done_stream));
    if (c->host_frames >= 1) {
        const int pslot = (int)((c->host_frames - 1) % kHostSlots);
        BMFR_CUDA_TRY(cudaEventRecord(c->frame_done[pslot], done_stream));
    }
    if (h_out) {  // bmfr.cpp:479-480
        BMFR_CUDA_TRY(cudaStreamWaitEvent(c->d2h_stream, c->frame_done[slot], 0));
        const size_t off = (size_t)(c->geo.own_y0 - c->geo.row0) * c->geo.width * 3;
        const size_t obytes = (size_t)(c->geo.own_y1 - c->geo.own_y0) * c->geo.width * 3 * sizeof(float);
        BMFR_CUDA_TRY(cudaMemcpyAsync(h_out + off, d_result + off, obytes, cudaMemcpyDeviceToHost, c->d2h_stream));
    }
    BMFR_CUDA_TRY(cudaEventRecord(c->d2h_done[rslot], c->d2h_stream));
    ++c->host_frames;
    return BMFR_OK;
}

int bmfr_join(bmfr_ctx* c) {
    if (!c) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_join: null context");
    if (!c->ov.on || c->seq == 0) return BMFR_OK;
    BMFR_CUDA_TRY(cudaSetDevice(c->prm.device));
    // P(f) is the last kernel of frame f, and the post passes run in order on one stream
    BMFR_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ov.e_p[c->last_parity()], 0));
    return BMFR_OK;
}

int bmfr_sync(bmfr_ctx* c) {
    if (!c) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_sync: null context");
    BMFR_CUDA_TRY(cudaSetDevice(c->prm.device));
    if (c->h2d_stream) BMFR_CUDA_TRY(cudaStreamSynchronize(c->h2d_stream));
    BMFR_CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (cudaStream_t st : {c->ov.s_r, c->ov.s_f, c->ov.s_p})
        if (st) BMFR_CUDA_TRY(cudaStreamSynchronize(st));
    if (c->d2h_stream) BMFR_CUDA_TRY(cudaStreamSynchronize(c->d2h_stream));
    unsigned int timed_out = 0;
    BMFR_CUDA_TRY(cudaMemcpy(&timed_out, c->d_flags + 2, sizeof(unsigned int), cudaMemcpyDeviceToHost));
    if (timed_out || c->failed) {  // sticky: the frame ran on stale halo rows and nothing was signalled to the neighbours
        c->failed = true;
        return bmfr_set_error(BMFR_ERR_SEQUENCE, "bmfr_sync: a neighbouring strip did not deliver its halo rows within %d ms (the contexts "
                                                 "must submit the same frames); this context has failed and refuses further frames",
                              c->prm.halo_timeout_ms > 0 ? c->prm.halo_timeout_ms : 10000);
    }
    int oob = 0;
    BMFR_CUDA_TRY(cudaMemcpy(&oob, c->d_oob, sizeof(int), cudaMemcpyDeviceToHost));
    if (oob) {
        cudaMemset(c->d_oob, 0, sizeof(int));
        return bmfr_set_error(BMFR_ERR_HALO_TOO_SMALL,
                              "bmfr_sync: a gather or block left rows [%d,%d) held by this strip (halo_rows=%d too small "
                              "for the camera motion)", c->geo.row0, c->geo.row1, c->prm.halo_rows);
    }
    return BMFR_OK;
}

// accum / result as the caller expects them (3 floats per pixel): the buffer itself, or — padded context — a compacted copy
// made on the context's stream (valid until the next call for the same buffer; device work is ordered by that stream).
__global__ void hist_compact_kernel(const float4* __restrict__ src, float* __restrict__ dst, size_t npix) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float4 v = src[i];
    dst[3 * i] = v.x; dst[3 * i + 1] = v.y; dst[3 * i + 2] = v.z;
}
static float* hist_view(bmfr_ctx* c, int which, float* buf, size_t npix) {
    if (c->hist_stride == 3) return buf;
    if (!c->hist_compact[which] && dev_alloc(&c->hist_compact[which], npix * 3, "compacted history") != 0) return nullptr;
    if (c->ov.on) bmfr_join(c);  // overlapped frames: the post pass that wrote it runs on its own stream
    hist_compact_kernel<<<(unsigned int)((npix + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const float4*>(buf), c->hist_compact[which], npix);
    return cudaGetLastError() == cudaSuccess ? c->hist_compact[which] : nullptr;
}

int bmfr_get_buffer(bmfr_ctx* c, int buffer, void** d_ptr, size_t* bytes) {
    if (!c || !d_ptr || !bytes) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_buffer: null argument");
    const size_t npix = rows_of(c) * (size_t)c->geo.width;
    const size_t nb = (size_t)c->geo.blocks_x * c->geo.blocks_y;
    int nf = BMFR_FEATURES, ns = BMFR_FEATURES_SCALED;
    bmfr_feature_counts(c->prm.feature_set, &nf, &ns);
    void* p = nullptr;
    size_t n = 0;
    const bool last_odd = c->last_parity() == 1;  // overlapped frames: which copy the last frame wrote
    // after the swap at the end of a frame the buffers that frame wrote are the "previous" halves
    switch (buffer) {
        case BMFR_BUF_NOISY_ACC: p = c->noisy_acc.previous(); n = npix * 12; break;
        case BMFR_BUF_SPP: p = c->spp.previous(); n = npix; break;
        case BMFR_BUF_PREV_PIXELS: p = last_odd ? c->ov.prev_pixels : c->prev_pixels; n = npix * 8; break;
        case BMFR_BUF_ACCEPT: p = last_odd ? c->ov.accept : c->accept; n = npix; break;
        case BMFR_BUF_TMP_DATA: p = c->tmp_data; n = (size_t)c->tmp_block_rows * c->geo.blocks_x * BMFR_BUFFER_COUNT * BMFR_BLOCK_PIXELS * 4; break;
        case BMFR_BUF_WEIGHTS: p = last_odd ? c->ov.weights : c->weights; n = nb * nf * 3 * 4; break;
        case BMFR_BUF_MINS_MAXS: p = last_odd ? c->ov.mins_maxs : c->mins_maxs; n = nb * ns * 2 * 4; break;
        case BMFR_BUF_FILTERED: p = c->filtered; n = npix * 12; break;
        case BMFR_BUF_ACCUM: p = hist_view(c, 0, c->accum.previous(), npix); n = npix * 12; break;
        case BMFR_BUF_TONE_MAPPED: p = c->tone_mapped; n = npix * 12; break;
        case BMFR_BUF_RESULT: p = hist_view(c, 1, c->result.previous(), npix); n = npix * 12; break;
        case BMFR_BUF_NOISE_TILE: p = last_odd ? c->ov.noise : c->noise; n = (size_t)(nf - 1) * BMFR_BLOCK_PIXELS * 8; break;
        default: return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_buffer: unknown buffer %d", buffer);
    }
    if (!p)
        return bmfr_set_error(BMFR_ERR_UNSUPPORTED, "bmfr_get_buffer: buffer %d is not materialised in this mode", buffer);
    *d_ptr = p;
    *bytes = n;
    return BMFR_OK;
}

int bmfr_read_buffer(bmfr_ctx* c, int buffer, void* h_dst, size_t bytes) {
    void* p;
    size_t n;
    int st = bmfr_get_buffer(c, buffer, &p, &n);
    if (st != 0) return st;
    if (!h_dst || bytes > n) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_read_buffer: %zu bytes requested, buffer has %zu", bytes, n);
    st = bmfr_sync(c);
    if (st != 0) return st;
    BMFR_CUDA_TRY(cudaMemcpy(h_dst, p, bytes, cudaMemcpyDeviceToHost));
    return BMFR_OK;
}

int bmfr_get_stage_ms(bmfr_ctx* c, int frame, float ms[BMFR_STAGE_COUNT]) {
    if (!c || !ms) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_stage_ms: null argument");
    if (c->prof.empty()) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_stage_ms: context created with profile=0");
    StageEvents& s = c->prof[(size_t)frame % kProfileSlots];
    if (!s.created || s.frame != frame) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_stage_ms: frame %d not recorded", frame);
    for (int i = 0; i < BMFR_STAGE_COUNT; ++i) ms[i] = 0.f;
    if (c->prm.mode == BMFR_MODE_STAGED) {
        BMFR_CUDA_TRY(cudaEventSynchronize(s.ev[5]));
        for (int i = 0; i < 5; ++i) BMFR_CUDA_TRY(cudaEventElapsedTime(&ms[i], s.ev[i], s.ev[i + 1]));
        BMFR_CUDA_TRY(cudaEventElapsedTime(&ms[BMFR_STAGE_TOTAL], s.ev[0], s.ev[5]));  // K1 start -> K5 end, bmfr.cpp:497-502
    } else {
        // FUSED: reproject -> ACCUM_NOISY, fit_qr -> FITTER, post -> TAA (it contains K3 and K4)
        BMFR_CUDA_TRY(cudaEventSynchronize(s.ev[3]));
        BMFR_CUDA_TRY(cudaEventElapsedTime(&ms[BMFR_STAGE_ACCUM_NOISY], s.ev[0], s.ev[1]));
        BMFR_CUDA_TRY(cudaEventElapsedTime(&ms[BMFR_STAGE_FITTER], s.ev[1], s.ev[2]));
        BMFR_CUDA_TRY(cudaEventElapsedTime(&ms[BMFR_STAGE_TAA], s.ev[2], s.ev[3]));
        BMFR_CUDA_TRY(cudaEventElapsedTime(&ms[BMFR_STAGE_TOTAL], s.ev[0], s.ev[3]));
    }
    return BMFR_OK;
}

int bmfr_get_fused_kernel_ms(bmfr_ctx* c, int frame, float ms[BMFR_FUSED_KERNEL_COUNT]) {
    if (!c || !ms) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_fused_kernel_ms: null argument");
    if (c->prof.empty() || c->prm.mode != BMFR_MODE_FUSED)
        return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_fused_kernel_ms: needs a FUSED context created with profile=1");
    StageEvents& s = c->prof[(size_t)frame % kProfileSlots];
    if (!s.created || s.frame != frame) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_fused_kernel_ms: frame %d not recorded", frame);
    BMFR_CUDA_TRY(cudaEventSynchronize(s.ev[BMFR_FUSED_KERNEL_COUNT]));
    for (int i = 0; i < BMFR_FUSED_KERNEL_COUNT; ++i) BMFR_CUDA_TRY(cudaEventElapsedTime(&ms[i], s.ev[i], s.ev[i + 1]));
    return BMFR_OK;
}

int bmfr_get_fused_kernel_busy_ms(bmfr_ctx* c, int frame, float ms[BMFR_FUSED_KERNEL_COUNT], float* frame_ms) {
    if (!c || !ms) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_fused_kernel_busy_ms: null argument");
    if (!c->d_stamps) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_fused_kernel_busy_ms: needs a FUSED context created with profile=2");
    int slot = -1;
    for (int i = 0; i < kProfileSlots; ++i)
        if (c->stamp_frame[i] == frame) slot = i;
    if (slot < 0) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_fused_kernel_busy_ms: frame %d not recorded", frame);
    int st = bmfr_sync(c);
    if (st != 0) return st;
    unsigned long long t[6];
    BMFR_CUDA_TRY(cudaMemcpy(t, c->d_stamps + (size_t)slot * 6, sizeof(t), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; ++k) ms[k] = (float)((double)(~t[2 * k + 1] - t[2 * k]) * 1e-6);
    if (frame_ms) *frame_ms = (float)((double)(~t[5] - t[0]) * 1e-6);
    return BMFR_OK;
}

int bmfr_get_fused_kernel_stamps(bmfr_ctx* c, int frame, unsigned long long ns[2 * BMFR_FUSED_KERNEL_COUNT]) {
    if (!c || !ns) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_fused_kernel_stamps: null argument");
    if (!c->d_stamps) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_fused_kernel_stamps: needs a FUSED context created with profile=2");
    int slot = -1;
    for (int i = 0; i < kProfileSlots; ++i)
        if (c->stamp_frame[i] == frame) slot = i;
    if (slot < 0) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_fused_kernel_stamps: frame %d not recorded", frame);
    int st = bmfr_sync(c);
    if (st != 0) return st;
    BMFR_CUDA_TRY(cudaMemcpy(ns, c->d_stamps + (size_t)slot * 6, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 3; ++k) ns[2 * k + 1] = ~ns[2 * k + 1];  // the end is kept as its complement (atomicMin)
    return BMFR_OK;
}

static int halo_check_neighbour(const bmfr_ctx* c, int side, int n_w, int n_h, int n_own_y0, int n_own_y1, int n_row0, int n_row1,
                                int n_swapped, long long n_seq) {
    const bmfr_geometry& g = c->geo;
    if (side != 0 && side != 1) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "halo connect: side must be 0 (above) or 1 (below)");
    if (n_w != g.width || n_h != g.height) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "halo connect: the neighbour denoises another image size");
    if ((side == 0 && n_own_y1 != g.own_y0) || (side == 1 && n_own_y0 != g.own_y1))
        return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "halo connect: strips [%d,%d) and [%d,%d) are not adjacent on side %d", g.own_y0,
                              g.own_y1, n_own_y0, n_own_y1, side);
    // every halo row must belong to the immediate neighbour
    if ((side == 0 && g.row0 < n_own_y0) || (side == 1 && g.row1 > n_own_y1) || (side == 0 && n_row1 > g.own_y1) ||
        (side == 1 && n_row0 < g.own_y0))
        return bmfr_set_error(BMFR_ERR_UNSUPPORTED, "halo connect: halo_rows exceed the neighbouring strip (strips must be at least halo_rows tall)");
    if (c->prm.mode != BMFR_MODE_FUSED)
        return bmfr_set_error(BMFR_ERR_UNSUPPORTED, "halo connect: the in-library halo exchange lives in the FUSED kernels (STAGED strips: "
                                                    "refresh the state rows yourself, bmfr_get_halo_plan)");
    // the flags count frames from zero and the halo rows of frames submitted earlier were never exchanged
    if (c->seq != 0 || n_seq != 0 || n_swapped != (c->noisy_acc.swapped ? 1 : 0))
        return bmfr_set_error(BMFR_ERR_SEQUENCE, "halo connect: contexts must be connected before their first frame");
    return BMFR_OK;
}

int bmfr_halo_export(bmfr_ctx* c, void* blob, size_t blob_bytes) {
    if (!c || !blob || blob_bytes < sizeof(HaloBlob)) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_halo_export: need %zu bytes", sizeof(HaloBlob));
    BMFR_CUDA_TRY(cudaSetDevice(c->prm.device));
    HaloBlob b;
    memset(&b, 0, sizeof(b));
    b.magic = kHaloMagic; b.device = c->prm.device; b.width = c->geo.width; b.height = c->geo.height;
    b.row0 = c->geo.row0; b.row1 = c->geo.row1; b.own_y0 = c->geo.own_y0; b.own_y1 = c->geo.own_y1;
    b.swapped = c->noisy_acc.swapped ? 1 : 0; b.seq = c->seq; b.mode = c->prm.mode;
    {
        cudaDeviceProp prop;
        BMFR_CUDA_TRY(cudaGetDeviceProperties(&prop, c->prm.device));
        memcpy(b.uuid, &prop.uuid, 16);
    }
    for (int i = 0; i < 2; ++i) {
        BMFR_CUDA_TRY(cudaIpcGetMemHandle(&b.noisy_acc[i], c->noisy_acc.buf[i]));
        BMFR_CUDA_TRY(cudaIpcGetMemHandle(&b.spp[i], c->spp.buf[i]));
        BMFR_CUDA_TRY(cudaIpcGetMemHandle(&b.accum[i], c->accum.buf[i]));
        BMFR_CUDA_TRY(cudaIpcGetMemHandle(&b.result[i], c->result.buf[i]));
    }
    BMFR_CUDA_TRY(cudaIpcGetMemHandle(&b.flags, c->d_flags));
    memcpy(blob, &b, sizeof(b));
    return BMFR_OK;
}

int bmfr_halo_connect(bmfr_ctx* c, int side, const void* neighbour_blob, size_t blob_bytes) {
    if (!c || !neighbour_blob || blob_bytes < sizeof(HaloBlob)) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_halo_connect: bad blob");
    HaloBlob b;
    memcpy(&b, neighbour_blob, sizeof(b));
    if (b.magic != kHaloMagic) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_halo_connect: not a bmfr_halo_export blob");
    int st = halo_check_neighbour(c, side, b.width, b.height, b.own_y0, b.own_y1, b.row0, b.row1, b.swapped, b.seq);
    if (st != 0) return st;
    if (b.mode != BMFR_MODE_FUSED) return bmfr_set_error(BMFR_ERR_UNSUPPORTED, "bmfr_halo_connect: the neighbour is not a FUSED context");
    BMFR_CUDA_TRY(cudaSetDevice(c->prm.device));
    bmfr_ctx::Peer& pr = c->peer[side];
    const unsigned int fl = cudaIpcMemLazyEnablePeerAccess;
    for (int i = 0; i < 2; ++i) {
        BMFR_CUDA_TRY(cudaIpcOpenMemHandle((void**)&pr.noisy_acc[i], b.noisy_acc[i], fl));
        BMFR_CUDA_TRY(cudaIpcOpenMemHandle((void**)&pr.spp[i], b.spp[i], fl));
        BMFR_CUDA_TRY(cudaIpcOpenMemHandle((void**)&pr.accum[i], b.accum[i], fl));
        BMFR_CUDA_TRY(cudaIpcOpenMemHandle((void**)&pr.result[i], b.result[i], fl));
    }
    BMFR_CUDA_TRY(cudaIpcOpenMemHandle((void**)&pr.flags, b.flags, fl));
    pr.row0 = b.row0; pr.row1 = b.row1; pr.own_y0 = b.own_y0; pr.own_y1 = b.own_y1;
    pr.ipc = true;
    {
        cudaDeviceProp prop;
        BMFR_CUDA_TRY(cudaGetDeviceProperties(&prop, c->prm.device));
        pr.same_device = memcmp(b.uuid, &prop.uuid, 16) == 0;
    }
    pr.connected = true;
    return BMFR_OK;
}

int bmfr_halo_connect_local(bmfr_ctx* c, int side, bmfr_ctx* n) {
    if (!c || !n || c == n) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_halo_connect_local: bad context");
    int st = halo_check_neighbour(c, side, n->geo.width, n->geo.height, n->geo.own_y0, n->geo.own_y1, n->geo.row0, n->geo.row1,
                                  n->noisy_acc.swapped ? 1 : 0, n->seq);
    if (st != 0) return st;
    if (n->prm.mode != BMFR_MODE_FUSED) return bmfr_set_error(BMFR_ERR_UNSUPPORTED, "bmfr_halo_connect_local: the neighbour is not a FUSED context");
    if (n->prm.device != c->prm.device) {
        BMFR_CUDA_TRY(cudaSetDevice(c->prm.device));
        int can = 0;
        BMFR_CUDA_TRY(cudaDeviceCanAccessPeer(&can, c->prm.device, n->prm.device));
        if (!can) return bmfr_set_error(BMFR_ERR_UNSUPPORTED, "bmfr_halo_connect_local: device %d cannot access device %d", c->prm.device, n->prm.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(n->prm.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return bmfr_check_cuda(e, "cudaDeviceEnablePeerAccess");
        cudaGetLastError();
    }
    bmfr_ctx::Peer& pr = c->peer[side];
    for (int i = 0; i < 2; ++i) {
        pr.noisy_acc[i] = n->noisy_acc.buf[i]; pr.spp[i] = n->spp.buf[i];
        pr.accum[i] = n->accum.buf[i]; pr.result[i] = n->result.buf[i];
    }
    pr.flags = n->d_flags;
    pr.row0 = n->geo.row0; pr.row1 = n->geo.row1; pr.own_y0 = n->geo.own_y0; pr.own_y1 = n->geo.own_y1;
    pr.ipc = false;
    pr.same_device = n->prm.device == c->prm.device;
    pr.connected = true;
    return BMFR_OK;
}

int bmfr_get_halo_plan(const bmfr_ctx* c, int side, bmfr_halo_plan* out) {
    if (!c || !out || (side != 0 && side != 1)) return bmfr_set_error(BMFR_ERR_INVALID_ARGUMENT, "bmfr_get_halo_plan: bad argument");
    const bmfr_geometry& g = c->geo;
    const int halo = c->prm.halo_rows;
    if (side == 0) {
        out->recv_y0 = g.row0; out->recv_y1 = g.own_y0;
        out->send_y0 = g.own_y0; out->send_y1 = (g.own_y0 > 0) ? ((g.own_y0 + halo < g.own_y1) ? g.own_y0 + halo : g.own_y1) : g.own_y0;
    } else {
        out->recv_y0 = g.own_y1; out->recv_y1 = g.row1;
        out->send_y1 = g.own_y1; out->send_y0 = (g.own_y1 < g.height) ? ((g.own_y1 - halo > g.own_y0) ? g.own_y1 - halo : g.own_y0) : g.own_y1;
    }
    return BMFR_OK;
}

}  // extern "C"
