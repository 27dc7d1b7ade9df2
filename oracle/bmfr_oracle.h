/* TEST INFRASTRUCTURE — not part of the product.
 *
 * C interface shared by the two CPU checkers of the BMFR hot path:
 *   oracle/libbmfr_oracle.so      — "port": plain-C restatement of /root/reference/opencl/bmfr.cl
 *                                   (oracle/bmfr_oracle.c)
 *   oracle/_ref/libbmfr_clref.so  — "reference": the reference's own bmfr.cl compiled as C++
 *                                   through the OpenCL-C shim in oracle/cl_shim/ (built only where
 *                                   /root/reference exists)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load these libraries.
 */
#ifndef BMFR_ORACLE_H
#define BMFR_ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_params {
    int width, height;            /* IMAGE_WIDTH / IMAGE_HEIGHT, bmfr.cpp:39-40 */
    double noise_amount;          /* NOISE_AMOUNT (double literal), bmfr.cpp:58 */
    float blend_alpha;            /* bmfr.cpp:60 */
    float second_blend_alpha;     /* bmfr.cpp:61 */
    float taa_blend_alpha;        /* bmfr.cpp:62 */
    float position_limit_squared; /* bmfr.cpp:226 */
    float normal_limit_squared;   /* bmfr.cpp:227 */
    int tmp_half;                 /* USE_HALF_PRECISION_IN_TMP_DATA, bmfr.cpp:88 */
    int keep_tmp;                 /* keep a copy of tmp_data as K1 left it (the fitter destroys it) */
    int k1_schedule;              /* 0: mirrored work-items of K1 run before in-image ones (every read
                                     of current_noisy sees the kernel's input — the intended semantics,
                                     SURVEY H2a); 1: plain row-major work-item order */
    int threads;                  /* OpenMP threads, 0 = runtime default */
} oracle_params;

/* same numbering as bmfr_buffer in include/bmfr_b200.h */
enum {
    ORACLE_BUF_NOISY_ACC = 0, ORACLE_BUF_SPP = 1, ORACLE_BUF_PREV_PIXELS = 2, ORACLE_BUF_ACCEPT = 3,
    ORACLE_BUF_TMP_DATA = 4, ORACLE_BUF_WEIGHTS = 5, ORACLE_BUF_MINS_MAXS = 6, ORACLE_BUF_FILTERED = 7,
    ORACLE_BUF_ACCUM = 8, ORACLE_BUF_TONE_MAPPED = 9, ORACLE_BUF_RESULT = 10, ORACLE_BUF_NOISE_TILE = 11
};

typedef struct oracle_state oracle_state;

const char* oracle_kind(void); /* "port" or "reference" */
oracle_state* oracle_create(const oracle_params* p);
void oracle_destroy(oracle_state* s);
/* One iteration of the frame loop bmfr.cpp:417-485 (uploads, five kernels, swap). */
int oracle_frame(oracle_state* s, int frame, const float* albedo, const float* normals, const float* positions,
                 const float* noisy, const float cam_prev[16], const float pixel_offset[2]);
/* Buffers as of the last frame, reference layouts; sizes follow the image (W*H pixels). */
const void* oracle_buffer(oracle_state* s, int id, size_t* bytes);
/* Wall-clock ms of the five kernels of the last frame + their sum. */
void oracle_stage_ms(oracle_state* s, double ms[6]);
/* random() of bmfr.cl:162-171 for known-answer tests. */
float oracle_random(unsigned int a);

#ifdef __cplusplus
}
#endif
#endif
