/* bmfr_b200 — C ABI of the B200-native BMFR per-frame denoise path.
 *
 * The reference (tcantenot/bmfr) has no plugin or FFI interface: the hot path is the body of the
 * frame loop in /root/reference/opencl/bmfr.cpp:417-485 plus the five OpenCL kernels it enqueues
 * (/root/reference/opencl/bmfr.cl:290,490,703,761,860).  This header is the boundary a maintainer
 * would bind instead of that loop body.  Every entry point cites the reference lines it replaces.
 *
 * Conventions
 *   - plain C, no exceptions across the boundary; every call returns BMFR_OK (0) or a negative
 *     bmfr_status, and bmfr_last_error() returns the message of the calling thread's last failure
 *     (the reference throws cl::Error and returns the CL code from main, bmfr.cpp:564-576);
 *   - one context = one CUDA device + one in-order stream (the reference's single in-order queue,
 *     bmfr.cpp:191).  A context is not thread-safe; distinct contexts are independent;
 *   - images are tightly packed interleaved RGB fp32, index (y*W + x)*3 + c (bmfr.cl:224-241);
 *   - there is no CPU fallback: without a CUDA device bmfr_create fails with BMFR_ERR_NO_DEVICE.
 */
#ifndef BMFR_B200_H
#define BMFR_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BMFR_B200_ABI_VERSION 5

/* Compile-time constants of the reference that are not tunable (bmfr.cpp:102-118). */
#define BMFR_BLOCK_EDGE 32          /* BLOCK_EDGE_LENGTH, bmfr.cpp:104 */
#define BMFR_BLOCK_PIXELS 1024      /* BLOCK_PIXELS, bmfr.cpp:105 */
#define BMFR_FEATURES 10            /* 1, n.xyz, p.xyz, p^2.xyz — bmfr.cpp:65-77 */
#define BMFR_FEATURES_NOT_SCALED 4  /* bmfr.cpp:65-69 */
#define BMFR_FEATURES_SCALED 6      /* bmfr.cpp:71-77 */
#define BMFR_BUFFER_COUNT 13        /* features + 3 colour channels, bmfr.cpp:202 */
#define BMFR_BLOCK_OFFSETS_COUNT 16 /* bmfr.cl:267 */

typedef enum bmfr_status {
    BMFR_OK = 0,
    BMFR_ERR_INVALID_ARGUMENT = -1,
    BMFR_ERR_NO_DEVICE = -2,      /* no CUDA device / driver: the product has no CPU path */
    BMFR_ERR_CUDA = -3,           /* a CUDA runtime call or kernel failed; see bmfr_last_error() */
    BMFR_ERR_OUT_OF_MEMORY = -4,
    BMFR_ERR_UNSUPPORTED = -5,    /* valid reference configuration that this build does not cover */
    BMFR_ERR_HALO_TOO_SMALL = -6, /* sharded context: a reprojection gather left strip + halo */
    BMFR_ERR_SEQUENCE = -7        /* frame > 0 submitted to a context that has no previous frame */
} bmfr_status;

/* Kernel structure.  STAGED launches the reference's five kernels one by one (bmfr.cpp:446-476).
 * FUSED launches three: reproject (accumulate_noisy_data per image pixel), fit_qr (the fitter as a
 * two-level QR straight from the per-pixel buffers) and post (weighted_sum +
 * accumulate_filtered_data + taa in one pass).  Both produce the same buffers the reference's loop
 * exposes to its caller; FUSED does not materialise tmp_data / filtered / tone_mapped in HBM. */
typedef enum bmfr_mode { BMFR_MODE_STAGED = 0, BMFR_MODE_FUSED = 1 } bmfr_mode;

/* Tuning defines of bmfr.cpp:56-62 and the two dataset thresholds of camera_matrices.h
 * (bmfr.cpp:226-227), as run-time parameters. */
typedef struct bmfr_params {
    int width;                    /* IMAGE_WIDTH,  bmfr.cpp:39 (full image, also when sharded) */
    int height;                   /* IMAGE_HEIGHT, bmfr.cpp:40 */
    int device;                   /* CUDA ordinal (PLATFORM_INDEX/DEVICE_INDEX, bmfr.cpp:33-34) */
    int mode;                     /* bmfr_mode */
    double noise_amount;          /* NOISE_AMOUNT — a double literal in the reference, bmfr.cpp:58 */
    float blend_alpha;            /* BLEND_ALPHA,        bmfr.cpp:60 */
    float second_blend_alpha;     /* SECOND_BLEND_ALPHA, bmfr.cpp:61 */
    float taa_blend_alpha;        /* TAA_BLEND_ALPHA,    bmfr.cpp:62 */
    float position_limit_squared; /* POSITION_LIMIT_SQUARED, bmfr.cpp:226 */
    float normal_limit_squared;   /* NORMAL_LIMIT_SQUARED,   bmfr.cpp:227 */
    int tmp_half;                 /* USE_HALF_PRECISION_IN_TMP_DATA, bmfr.cpp:88 (the reference ships 1; the default here
                                     is 0 = fp32 fitter).  1 needs mode = STAGED and implies reference_order = 1 */
    int profile;                  /* 1: record per-stage CUDA-event times (bmfr.cpp:386-397); 2 (FUSED): no events between the
                                     kernels (their programmatic chaining stays intact), every kernel stamps the start of its
                                     first and the end of its last CTA instead — bmfr_get_fused_kernel_busy_ms */
    /* Strip sharding (no counterpart in the reference, which is single-device).  The context owns
     * image rows [strip_y0, strip_y1) and stores rows [strip_y0 - halo_rows, strip_y1 + halo_rows)
     * clipped to the image.  strip_y0 = strip_y1 = 0 means the whole image, halo_rows ignored. */
    int strip_y0;
    int strip_y1;
    int halo_rows;
    void* stream;                 /* cudaStream_t to enqueue on; NULL: the context creates one */
    /* STAGED only.  1: the fitter and the weighted sum run in the reference's own operation order
     * (256-thread groups, the reduction trees of bmfr.cl:26-87, 13 Householder columns, division in
     * scale()), which makes EVERY buffer of the loop bit-identical to the reference kernels' arithmetic
     * — the compatibility mode; slower than the default fitter. */
    int reference_order;
    /* FUSED contexts without profiling only (ignored otherwise).  1: consecutive frames may overlap on the
     * device.  The reprojection of frame f+1 does not depend on the fit and the post pass of
     * frame f, so the three kernels are enqueued on three internal streams linked by events (per-frame
     * temporaries double-buffered) and the tail of one kernel is filled by the next frame's work.  The
     * context's stream is then ordered BEFORE a frame's kernels (the caller's producers of the inputs) but
     * not after them: the output (d_out, bmfr_get_buffer) is valid, and the input buffers of the last two
     * frames may be overwritten, only after bmfr_sync() or, on the stream, after bmfr_join().  bmfr_denoise_frame_host handles that ordering
     * itself.  Strip contexts may mix both settings (the halo protocol is the same); a caller-driven exchange of the
     * state rows (which needs the in-order stream) cannot be combined with overlap_frames = 1.
     * 0 (default): one in-order stream, the reference's queue semantics (bmfr.cpp:191). */
    int overlap_frames;
    /* FUSED only: how the per-block least-squares problem of the fitter (bmfr.cl:546-699) is solved.
     *   BMFR_FIT_GRAM (0, default): centred normal equations — one 13x13 Gram matrix per block accumulated in fp32,
     *       reduced once, Cholesky + substitutions in fp64;
     *   BMFR_FIT_TSQR (1): two-level Householder/MGS QR in fp32 registers (the round-1 kernel; ten dependent warp
     *       reductions per block).
     * Both are held to the same colour tolerance against the reference's Householder QR. */
    int fit_method;
    /* Strip contexts with a connected neighbour: how long a kernel waits on the device for the neighbour's halo rows
     * before it gives up (the context then fails: every later call returns BMFR_ERR_SEQUENCE).  0 = 10000 ms. */
    int halo_timeout_ms;
    /* bmfr_feature_set.  Lists other than the default need mode = FUSED and fit_method = BMFR_FIT_GRAM. */
    int feature_set;
} bmfr_params;

enum { BMFR_FIT_GRAM = 0, BMFR_FIT_TSQR = 1 };

/* Feature lists.  In the reference the list is a compile-time string (NOT_SCALED_FEATURE_BUFFERS / SCALED_FEATURE_BUFFERS,
 * bmfr.cpp:63-77; "if you want to use other than normal and world_position data you have to make it available in the
 * first accumulation kernel and in the weighted sum kernel", :63-64); here the FUSED kernels are instantiated for these
 * lists over the same two inputs and bmfr_params.feature_set picks one at run time.  BMFR_FEATURES / _SCALED /
 * BMFR_BUFFER_COUNT above are the sizes of the default list and the maxima. */
typedef enum bmfr_feature_set {
    BMFR_FEATURE_SET_DEFAULT = 0,  /* 1, normal.xyz | world_position.xyz, world_position.xyz^2  (10 features, 6 scaled: bmfr.cpp:65-77) */
    BMFR_FEATURE_SET_LINEAR = 1,   /* 1, normal.xyz | world_position.xyz                         ( 7 features, 3 scaled) */
    BMFR_FEATURE_SET_POSITION = 2, /* 1             | world_position.xyz, world_position.xyz^2  ( 7 features, 6 scaled) */
    BMFR_FEATURE_SET_COUNT_ = 3
} bmfr_feature_set;
/* Feature count and scaled-feature count of a list (what FEATURES_NOT_SCALED + FEATURES_SCALED / FEATURES_SCALED are in
 * bmfr.cpp:195-199); sizes of BMFR_BUF_WEIGHTS (NB*F*3), BMFR_BUF_MINS_MAXS (NB*S*2), BMFR_BUF_NOISE_TILE ((F-1)*1024). */
int bmfr_feature_counts(int feature_set, int* features, int* scaled);

typedef struct bmfr_ctx bmfr_ctx;

/* Buffers of the reference's host loop (bmfr.cpp:315-343) that can be inspected.  All are returned
 * in the reference's own layout. */
typedef enum bmfr_buffer {
    BMFR_BUF_NOISY_ACC = 0,   /* float[rows*W*3]: current_noisy after K1's in-place store, bmfr.cl:481 */
    BMFR_BUF_SPP = 1,         /* uchar[rows*W]:   current_spp, bmfr.cl:442 */
    BMFR_BUF_PREV_PIXELS = 2, /* float2[rows*W]:  out_prev_frame_pixel, bmfr.cl:482 */
    BMFR_BUF_ACCEPT = 3,      /* uchar[rows*W]:   accept_bools, bmfr.cl:483 */
    BMFR_BUF_TMP_DATA = 4,    /* float[NB*13*1024]: K1's block-planar output, bmfr.cl:455-476 (STAGED only) */
    BMFR_BUF_WEIGHTS = 5,     /* float[NB*10*3]:  bmfr.cl:694-699 */
    BMFR_BUF_MINS_MAXS = 6,   /* float[NB*6*2]:   bmfr.cl:531-535 */
    BMFR_BUF_FILTERED = 7,    /* float[rows*W*3]: weighted_sum output, bmfr.cl:757 (STAGED only) */
    BMFR_BUF_ACCUM = 8,       /* float[rows*W*3]: accumulated_frame, bmfr.cl:849 */
    BMFR_BUF_TONE_MAPPED = 9, /* float[rows*W*3]: tone_mapped_frame, bmfr.cl:856 (STAGED only) */
    BMFR_BUF_RESULT = 10,     /* float[rows*W*3]: result_frame, bmfr.cl:973 */
    BMFR_BUF_NOISE_TILE = 11, /* double[9*1024]: this frame's add_random() terms, bmfr.cl:173-182 */
    BMFR_BUF_COUNT_ = 12
} bmfr_buffer;

/* Stage indices for bmfr_get_stage_ms, in the order the reference prints them (bmfr.cpp:399-412). */
enum {
    BMFR_STAGE_ACCUM_NOISY = 0,
    BMFR_STAGE_FITTER = 1,
    BMFR_STAGE_WEIGHTED_SUM = 2,
    BMFR_STAGE_ACCUM_FILTERED = 3,
    BMFR_STAGE_TAA = 4,
    BMFR_STAGE_TOTAL = 5,
    BMFR_STAGE_COUNT = 6
};

/* Geometry derived the way bmfr.cpp:107-118 derives it. */
typedef struct bmfr_geometry {
    int width, height;
    int workset_width, workset_height;   /* WORKSET_WIDTH/HEIGHT, bmfr.cpp:107-110 */
    int margin_width, margin_height;     /* WORKSET_WITH_MARGINS_*, bmfr.cpp:111-112 */
    int blocks_x, blocks_y;              /* margin grid in 32x32 blocks, bmfr.cpp:117-118 */
    int row0, row1;                      /* image rows held by this context (strip + halo) */
    int own_y0, own_y1;                  /* image rows owned (written to the output) */
    int block_row0, block_row1;          /* widest block-row range a frame can touch on this strip */
} bmfr_geometry;

const char* bmfr_last_error(void);
int bmfr_abi_version(void);

/* Fills *p with the reference's defaults (bmfr.cpp:56-62,88 with tmp_half forced to 0) for a
 * whole-image, FUSED, single-GPU context. */
void bmfr_default_params(bmfr_params* p, int width, int height);

/* BLOCK_OFFSETS[frame % 16], bmfr.cl:267-285. */
void bmfr_block_offset(int frame, int* off_x, int* off_y);

/* Replaces buffer creation + static argument binding, bmfr.cpp:315-384. */
int bmfr_create(const bmfr_params* params, bmfr_ctx** out_ctx);
void bmfr_destroy(bmfr_ctx* ctx);
int bmfr_get_geometry(const bmfr_ctx* ctx, bmfr_geometry* out);

/* One iteration of the frame loop, bmfr.cpp:429-476 + the swap at :483-484, with DEVICE pointers.
 *   d_albedo, d_normal, d_position, d_noisy : rows [row0,row1) of this frame's inputs, read-only.
 *       (The reference's K1 overwrites current_noisy in place, bmfr.cl:481; here the accumulated
 *       colour lives in an internal double buffer, BMFR_BUF_NOISY_ACC.)
 *   d_normal / d_position of frame f are read again while frame f+1 is processed (they are the
 *       "previous" halves of the reference's double buffers, bmfr.cpp:431-434): the caller keeps
 *       them alive and unchanged until the call for frame f+1 has completed.
 *   cam_prev      : camera_matrices[frame-1] (frame 0: ignored), 16 floats, bmfr.cpp:440-442
 *   pixel_offset  : pixel_offsets[frame], 2 floats, bmfr.cpp:443-444
 *   d_out         : receives result_buffer.current() rows [own_y0,own_y1) addressed like the
 *                   inputs (row0-relative), bmfr.cpp:479-480; NULL keeps it in BMFR_BUF_RESULT only.
 * Asynchronous: returns after enqueueing on the context's stream. */
int bmfr_denoise_frame(bmfr_ctx* ctx, int frame, const float* d_albedo, const float* d_normal,
                       const float* d_position, const float* d_noisy, const float cam_prev[16],
                       const float pixel_offset[2], float* d_out);

/* The same iteration with HOST pointers: the reference's four enqueueWriteBuffer calls
 * (bmfr.cpp:420-427) and its enqueueReadBuffer (bmfr.cpp:479-480) are part of the call.  Uploads
 * go to context-owned double buffers and overlap the previous frame's kernels; h_out (may be
 * NULL) is valid after bmfr_sync().  Host buffers should be page-locked for the copies to be
 * asynchronous. */
int bmfr_denoise_frame_host(bmfr_ctx* ctx, int frame, const float* h_albedo, const float* h_normal,
                            const float* h_position, const float* h_noisy, const float cam_prev[16],
                            const float pixel_offset[2], float* h_out);

/* queue.finish(), bmfr.cpp:486.  Also reports BMFR_ERR_HALO_TOO_SMALL for sharded contexts. */
int bmfr_sync(bmfr_ctx* ctx);

/* Orders the context's stream after every frame submitted so far, without blocking the host: what follows
 * on that stream (an event record, a copy of d_out, the producer of the next inputs) sees the frames
 * complete.  Only contexts with overlap_frames = 1 need it (their kernels run on internal streams); on
 * the others it does nothing, the stream already is in order. */
int bmfr_join(bmfr_ctx* ctx);

/* Device pointer + size of one of the loop's buffers as of the last submitted frame. */
int bmfr_get_buffer(bmfr_ctx* ctx, int buffer, void** d_ptr, size_t* bytes);
/* Synchronises and copies it to host memory. */
int bmfr_read_buffer(bmfr_ctx* ctx, int buffer, void* h_dst, size_t bytes);

/* Per-stage device time of one frame in ms (params.profile = 1), the quantity the reference
 * collects with OpenCL events (bmfr.cpp:488-506).  FUSED reports reproject in
 * BMFR_STAGE_ACCUM_NOISY, fit_qr in BMFR_STAGE_FITTER, post in BMFR_STAGE_TAA and zero
 * elsewhere. */
int bmfr_get_stage_ms(bmfr_ctx* ctx, int frame, float ms[BMFR_STAGE_COUNT]);
/* FUSED contexts: device time of each of the three kernels of one frame, in launch order. */
enum { BMFR_FUSED_REPROJECT = 0, BMFR_FUSED_FIT_QR = 1, BMFR_FUSED_POST = 2, BMFR_FUSED_KERNEL_COUNT = 3 };
int bmfr_get_fused_kernel_ms(bmfr_ctx* ctx, int frame, float ms[BMFR_FUSED_KERNEL_COUNT]);
/* FUSED contexts created with profile = 2: for each of the three kernels of one frame, the time from the start of its
 * first CTA to the end of its last one (device globaltimer), measured INSIDE the chained run — unlike the event times
 * above it contains no launch gap and the kernels may overlap; frame_ms (may be NULL): first CTA of the reprojection to
 * last CTA of the post pass. */
int bmfr_get_fused_kernel_busy_ms(bmfr_ctx* ctx, int frame, float ms[BMFR_FUSED_KERNEL_COUNT], float* frame_ms);
/* The stamps behind it: ns[2k] = start of the first CTA, ns[2k + 1] = end of the last CTA of kernel k of that frame, in
 * nanoseconds of the device's globaltimer (one clock per GPU: comparable between the frames and contexts of one device,
 * not between devices).  With them a caller draws the timeline of a chained / overlapped run — the stand-in for the
 * per-kernel CL_PROFILING_COMMAND_START / _END pairs the reference's GPUTimer reads (CLUtils.hpp:296-311). */
int bmfr_get_fused_kernel_stamps(bmfr_ctx* ctx, int frame, unsigned long long ns[2 * BMFR_FUSED_KERNEL_COUNT]);
/* Number of kernels this library has launched on the context so far. */
long long bmfr_kernel_launches(const bmfr_ctx* ctx);

/* Sharded contexts: state rows a neighbour must supply before the next frame.  For
 * side 0 (above) / 1 (below) reports, for the three temporal state buffers, the row ranges to
 * SEND (owned rows the neighbour's halo mirrors) and to RECEIVE (this context's halo rows). */
typedef struct bmfr_halo_plan {
    int send_y0, send_y1; /* image rows of this strip the neighbour needs */
    int recv_y0, recv_y1; /* image rows of the neighbour this strip needs */
} bmfr_halo_plan;
int bmfr_get_halo_plan(const bmfr_ctx* ctx, int side, bmfr_halo_plan* out);

/* Peer-to-peer halo exchange (option A; no counterpart in the single-device reference).  FUSED contexts
 * only.  Once a sharded context is connected to the contexts that own the strips above (side 0) and below (side 1),
 * the kernels of every bmfr_denoise_frame call do the exchange themselves: the CTAs next to a strip edge store the rows
 * a neighbour mirrors a second time, straight into the neighbour's halo rows (peer memory over NVLink), the last of
 * them raises a flag in the neighbour's memory, and they poll their own flags before they gather from halo rows — no
 * copy, signal or wait launches, no host round trip, no collective.  All contexts must submit the same frame sequence
 * and must be connected before their first frame.  A neighbour that does not deliver within params.halo_timeout_ms
 * fails the context: bmfr_sync reports BMFR_ERR_SEQUENCE and later frames are refused.
 *   bmfr_halo_export / bmfr_halo_connect : contexts in different processes (one rank per GPU): export an
 *       opaque blob (CUDA IPC handles + geometry), ship it to the neighbour with any transport, connect.
 *   bmfr_halo_connect_local              : both contexts live in this process. */
#define BMFR_HALO_BLOB_BYTES 1024
int bmfr_halo_export(bmfr_ctx* ctx, void* blob, size_t blob_bytes);
int bmfr_halo_connect(bmfr_ctx* ctx, int side, const void* neighbour_blob, size_t blob_bytes);
int bmfr_halo_connect_local(bmfr_ctx* ctx, int side, bmfr_ctx* neighbour);

/* ---- synthetic inputs ("synth-v1"): stands in for the dataset of bmfr.cpp:44-53 ---- */
void bmfr_synth_camera(int frame, int width, int height, int jitter, float cam_matrix[16], float pixel_offset[2]);
void bmfr_synth_limits(float* position_limit_squared, float* normal_limit_squared);
int bmfr_synth_frame_host(int width, int height, int y0, int y1, int frame, unsigned seed, float* albedo,
                          float* normal, float* position, float* noisy);
int bmfr_synth_frame_device(int width, int height, int y0, int y1, int frame, unsigned seed, float* d_albedo,
                            float* d_normal, float* d_position, float* d_noisy, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BMFR_B200_H */
