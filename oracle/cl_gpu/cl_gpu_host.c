/* TEST INFRASTRUCTURE — not part of the product.
 *
 * The reference's own, UNMODIFIED bmfr.cl on an OpenCL device of the box (on the B200 box: NVIDIA's
 * OpenCL ICD), behind the C interface of oracle/bmfr_oracle.h ("opencl" kind).  It is the second
 * checker of the parity tests and the "reference kernels on the same GPU" arm of bench.py
 * (SURVEY.md 7 step 0).
 *
 * The host side restates /root/reference/opencl/bmfr.cpp: build options :205-232, NDRanges :245-249,
 * buffers :315-343, static kernel arguments :349-383, the frame loop :417-485 (blocking uploads, five
 * launches with one profiling event each, double-buffer swap), profiling :488-506.
 *
 * There are no OpenCL headers in the image, so the few entry points used are declared by hand (their
 * C ABI is fixed by the OpenCL specification) and resolved with dlopen() from the ICD loader
 * (libOpenCL.so.1, which ships with the CUDA toolkit).  The loader finds no /etc/OpenCL/vendors on
 * the box; OCL_ICD_FILENAMES (which the loader honours, see `strings libOpenCL.so.1`) points it at
 * libnvidia-opencl.so.1 instead — nothing outside the repository is written.
 *
 * The kernel source and the feature-list build options are NOT in the repository: build_oracle.py
 * reads them from /root/reference at build time and compiles them into oracle/_ref/libbmfr_clgpu.so
 * (git-ignored) through the generated header included below.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../bmfr_oracle.h"
#include "bmfr_cl_source.gen.h" /* kBmfrClSource[], kBmfrClSourceLen, BMFR_CL_FEATURE_OPTIONS, BMFR_CL_* ints */

/* ---- the part of the OpenCL 1.2 C API used here (CL/cl.h is not in the image) ---- */
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef cl_ulong cl_bitfield;
typedef struct _cl_platform_id* cl_platform_id;
typedef struct _cl_device_id* cl_device_id;
typedef struct _cl_context* cl_context;
typedef struct _cl_command_queue* cl_command_queue;
typedef struct _cl_mem* cl_mem;
typedef struct _cl_program* cl_program;
typedef struct _cl_kernel* cl_kernel;
typedef struct _cl_event* cl_event;
#define CL_SUCCESS 0
#define CL_DEVICE_TYPE_GPU (1ull << 2)
#define CL_DEVICE_TYPE_ALL 0xFFFFFFFFull
#define CL_DEVICE_NAME 0x102B
#define CL_QUEUE_PROFILING_ENABLE (1ull << 1)
#define CL_MEM_READ_WRITE (1ull << 0)
#define CL_MEM_READ_ONLY (1ull << 2)
#define CL_PROGRAM_BUILD_LOG 0x1183
#define CL_PROFILING_COMMAND_START 0x1282
#define CL_PROFILING_COMMAND_END 0x1283

static struct {
    void* lib;
    cl_int (*GetPlatformIDs)(cl_uint, cl_platform_id*, cl_uint*);
    cl_int (*GetDeviceIDs)(cl_platform_id, cl_bitfield, cl_uint, cl_device_id*, cl_uint*);
    cl_int (*GetDeviceInfo)(cl_device_id, cl_uint, size_t, void*, size_t*);
    cl_context (*CreateContext)(const intptr_t*, cl_uint, const cl_device_id*, void*, void*, cl_int*);
    cl_command_queue (*CreateCommandQueue)(cl_context, cl_device_id, cl_bitfield, cl_int*);
    cl_program (*CreateProgramWithSource)(cl_context, cl_uint, const char**, const size_t*, cl_int*);
    cl_int (*BuildProgram)(cl_program, cl_uint, const cl_device_id*, const char*, void*, void*);
    cl_int (*GetProgramBuildInfo)(cl_program, cl_device_id, cl_uint, size_t, void*, size_t*);
    cl_kernel (*CreateKernel)(cl_program, const char*, cl_int*);
    cl_int (*SetKernelArg)(cl_kernel, cl_uint, size_t, const void*);
    cl_mem (*CreateBuffer)(cl_context, cl_bitfield, size_t, void*, cl_int*);
    cl_int (*EnqueueWriteBuffer)(cl_command_queue, cl_mem, cl_uint, size_t, size_t, const void*, cl_uint, const cl_event*, cl_event*);
    cl_int (*EnqueueReadBuffer)(cl_command_queue, cl_mem, cl_uint, size_t, size_t, void*, cl_uint, const cl_event*, cl_event*);
    cl_int (*EnqueueNDRangeKernel)(cl_command_queue, cl_kernel, cl_uint, const size_t*, const size_t*, const size_t*, cl_uint,
                                   const cl_event*, cl_event*);
    cl_int (*EnqueueFillBuffer)(cl_command_queue, cl_mem, const void*, size_t, size_t, size_t, cl_uint, const cl_event*, cl_event*);
    cl_int (*Finish)(cl_command_queue);
    cl_int (*GetEventProfilingInfo)(cl_event, cl_uint, size_t, void*, size_t*);
    cl_int (*ReleaseEvent)(cl_event);
    cl_int (*ReleaseMemObject)(cl_mem);
    cl_int (*ReleaseKernel)(cl_kernel);
    cl_int (*ReleaseProgram)(cl_program);
    cl_int (*ReleaseCommandQueue)(cl_command_queue);
    cl_int (*ReleaseContext)(cl_context);
} cl;

static char g_error[4096] = "";
const char* oracle_last_error(void) { return g_error; }

static int load_cl(void) {
    if (cl.lib) return 0;
    /* the loader reads these when the first platform query is made */
    if (!getenv("OCL_ICD_FILENAMES") && !getenv("OCL_ICD_VENDORS")) setenv("OCL_ICD_FILENAMES", "libnvidia-opencl.so.1", 0);
    const char* names[] = {getenv("BMFR_OPENCL_LIB"), "libOpenCL.so.1", "/usr/local/cuda/lib64/libOpenCL.so.1", "libOpenCL.so"};
    for (size_t i = 0; i < sizeof(names) / sizeof(names[0]) && !cl.lib; ++i)
        if (names[i]) cl.lib = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
    if (!cl.lib) {
        snprintf(g_error, sizeof(g_error), "no OpenCL ICD loader (libOpenCL.so.1): %s", dlerror());
        return -1;
    }
#define SYM(name)                                                              \
    do {                                                                       \
        *(void**)(&cl.name) = dlsym(cl.lib, "cl" #name);                       \
        if (!cl.name) {                                                        \
            snprintf(g_error, sizeof(g_error), "cl" #name " not exported");    \
            return -1;                                                         \
        }                                                                      \
    } while (0)
    SYM(GetPlatformIDs); SYM(GetDeviceIDs); SYM(GetDeviceInfo); SYM(CreateContext); SYM(CreateCommandQueue);
    SYM(CreateProgramWithSource); SYM(BuildProgram); SYM(GetProgramBuildInfo); SYM(CreateKernel); SYM(SetKernelArg);
    SYM(CreateBuffer); SYM(EnqueueWriteBuffer); SYM(EnqueueReadBuffer); SYM(EnqueueNDRangeKernel); SYM(EnqueueFillBuffer);
    SYM(Finish); SYM(GetEventProfilingInfo); SYM(ReleaseEvent); SYM(ReleaseMemObject); SYM(ReleaseKernel); SYM(ReleaseProgram);
    SYM(ReleaseCommandQueue); SYM(ReleaseContext);
#undef SYM
    return 0;
}

/* Double_buffer<cl::Buffer>, bmfr.cpp:122-135: current() is b until the first swap */
typedef struct {
    cl_mem a, b;
    int swapped;
} dbuf;
static cl_mem db_cur(const dbuf* d) { return d->swapped ? d->a : d->b; }
static cl_mem db_prev(const dbuf* d) { return d->swapped ? d->b : d->a; }

enum { K_ACCUM_NOISY, K_FITTER, K_WEIGHTED_SUM, K_ACCUM_FILTERED, K_TAA, K_COUNT };
static const char* kKernelNames[K_COUNT] = {"accumulate_noisy_data", "fitter", "weighted_sum", "accumulate_filtered_data", "taa"};

struct oracle_state {
    oracle_params p;
    int W, H, Ww, Hw, Wm, Hm, NB;
    cl_device_id dev;
    cl_context ctx;
    cl_command_queue q;
    cl_program prog;
    cl_kernel k[K_COUNT];
    dbuf normals, positions, noisy, out, result, spp;
    cl_mem in_buffer, filtered, prev_pixels, accept, albedo, tone_mapped, weights, mins_maxs;
    void* host[12];
    size_t host_bytes[12];
    double ms[6];
    char device_name[256];
};

const char* oracle_kind(void) { return "opencl"; }
const char* oracle_device_name(oracle_state* s) { return s ? s->device_name : ""; }

/* random() of bmfr.cl:162-171 restated for the known-answer interface (pure integer arithmetic) */
float oracle_random(unsigned int a) {
    a = (a + 0x7ed55d16u) + (a << 12);
    a = (a ^ 0xc761c23cu) ^ (a >> 19);
    a = (a + 0x165667b1u) + (a << 5);
    a = (a + 0xd3a2646cu) ^ (a << 9);
    a = (a + 0xfd7046c5u) + (a << 3);
    a = (a ^ 0xb55a4f09u) ^ (a >> 16);
    return (float)a / 4294967296.0f;
}

#define CL_TRY(call, what)                                                                   \
    do {                                                                                     \
        cl_int _e = (call);                                                                  \
        if (_e != CL_SUCCESS) {                                                              \
            snprintf(g_error, sizeof(g_error), "%s failed with OpenCL error %d", what, _e);  \
            goto fail;                                                                       \
        }                                                                                    \
    } while (0)

static cl_mem make_buffer(oracle_state* s, cl_bitfield flags, size_t bytes, cl_int* err) {
    cl_mem m = cl.CreateBuffer(s->ctx, flags, bytes, NULL, err);
    if (m && *err == CL_SUCCESS) {  /* defined contents, like the zero-initialised host vectors of the CPU checkers */
        const unsigned char zero = 0;
        *err = cl.EnqueueFillBuffer(s->q, m, &zero, 1, 0, bytes, 0, NULL, NULL);
    }
    return m;
}

void oracle_destroy(oracle_state* s) {
    if (!s) return;
    if (s->q) cl.Finish(s->q);
    for (int i = 0; i < K_COUNT; ++i)
        if (s->k[i]) cl.ReleaseKernel(s->k[i]);
    cl_mem all[] = {s->normals.a, s->normals.b, s->positions.a, s->positions.b, s->noisy.a, s->noisy.b, s->out.a, s->out.b,
                    s->result.a, s->result.b, s->spp.a, s->spp.b, s->in_buffer, s->filtered, s->prev_pixels, s->accept,
                    s->albedo, s->tone_mapped, s->weights, s->mins_maxs};
    for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); ++i)
        if (all[i]) cl.ReleaseMemObject(all[i]);
    if (s->prog) cl.ReleaseProgram(s->prog);
    if (s->q) cl.ReleaseCommandQueue(s->q);
    if (s->ctx) cl.ReleaseContext(s->ctx);
    for (int i = 0; i < 12; ++i) free(s->host[i]);
    free(s);
}

oracle_state* oracle_create(const oracle_params* p) {
    if (!p || p->width < 32 || p->height < 32) return NULL;
    if (load_cl() != 0) return NULL;
    oracle_state* s = (oracle_state*)calloc(1, sizeof(*s));
    if (!s) return NULL;
    s->p = *p;
    s->W = p->width; s->H = p->height;
    s->Ww = 32 * ((s->W + 31) / 32); s->Hw = 32 * ((s->H + 31) / 32);  /* bmfr.cpp:107-112 */
    s->Wm = s->Ww + 32; s->Hm = s->Hw + 32;
    s->NB = (s->Wm / 32) * (s->Hm / 32);
    cl_int err = 0;
    cl_platform_id plats[8];
    cl_uint nplat = 0, ndev = 0;
    CL_TRY(cl.GetPlatformIDs(8, plats, &nplat), "clGetPlatformIDs");
    if (nplat == 0) { snprintf(g_error, sizeof(g_error), "no OpenCL platform"); goto fail; }
    /* PLATFORM_INDEX 0 / DEVICE_INDEX 0 (bmfr.cpp:33-34), preferring a GPU */
    for (cl_uint i = 0; i < nplat && ndev == 0; ++i)
        if (cl.GetDeviceIDs(plats[i], CL_DEVICE_TYPE_GPU, 1, &s->dev, &ndev) != CL_SUCCESS) ndev = 0;
    for (cl_uint i = 0; i < nplat && ndev == 0; ++i)
        if (cl.GetDeviceIDs(plats[i], CL_DEVICE_TYPE_ALL, 1, &s->dev, &ndev) != CL_SUCCESS) ndev = 0;
    if (ndev == 0) { snprintf(g_error, sizeof(g_error), "no OpenCL device"); goto fail; }
    cl.GetDeviceInfo(s->dev, CL_DEVICE_NAME, sizeof(s->device_name) - 1, s->device_name, NULL);
    s->ctx = cl.CreateContext(NULL, 1, &s->dev, NULL, NULL, &err);
    CL_TRY(err, "clCreateContext");
    s->q = cl.CreateCommandQueue(s->ctx, s->dev, CL_QUEUE_PROFILING_ENABLE, &err);  /* in-order, bmfr.cpp:191 */
    CL_TRY(err, "clCreateCommandQueue");

    /* build options, bmfr.cpp:205-232.  The limits are streamed with nine significant digits instead of
     * ostream's default six, so that the float the CPU checkers use is the one the kernel sees. */
    {
        char opts[4096];
        snprintf(opts, sizeof(opts),
                 " -D BUFFER_COUNT=%d -D FEATURES_NOT_SCALED=%d -D FEATURES_SCALED=%d -D IMAGE_WIDTH=%d -D IMAGE_HEIGHT=%d"
                 " -D WORKSET_WIDTH=%d -D WORKSET_HEIGHT=%d -D FEATURE_BUFFERS=%s -D LOCAL_WIDTH=%d -D LOCAL_HEIGHT=%d"
                 " -D WORKSET_WITH_MARGINS_WIDTH=%d -D WORKSET_WITH_MARGINS_HEIGHT=%d -D BLOCK_EDGE_LENGTH=%d -D BLOCK_PIXELS=%d"
                 " -D R_EDGE=%d -D NOISE_AMOUNT=%.17g -D BLEND_ALPHA=%.9gf -D SECOND_BLEND_ALPHA=%.9gf -D TAA_BLEND_ALPHA=%.9gf"
                 " -D POSITION_LIMIT_SQUARED=%.9g -D NORMAL_LIMIT_SQUARED=%.9g -D COMPRESSED_R=%d -D CACHE_TMP_DATA=%d"
                 " -D ADD_REQD_WG_SIZE=%d -D LOCAL_SIZE=%d -D USE_HALF_PRECISION_IN_TMP_DATA=%d",
                 BMFR_CL_BUFFER_COUNT, BMFR_CL_FEATURES_NOT_SCALED, BMFR_CL_FEATURES_SCALED, s->W, s->H, s->Ww, s->Hw,
                 BMFR_CL_FEATURE_OPTIONS, BMFR_CL_LOCAL_WIDTH, BMFR_CL_LOCAL_HEIGHT, s->Wm, s->Hm, 32, 1024, BMFR_CL_BUFFER_COUNT - 2,
                 p->noise_amount, (double)p->blend_alpha, (double)p->second_blend_alpha, (double)p->taa_blend_alpha,
                 (double)p->position_limit_squared, (double)p->normal_limit_squared, BMFR_CL_COMPRESSED_R, BMFR_CL_CACHE_TMP_DATA,
                 BMFR_CL_ADD_REQD_WG_SIZE, BMFR_CL_LOCAL_SIZE, p->tmp_half ? 1 : 0);
        /* BMFR_OPENCL_STRICT_FP=1: the same source under the arithmetic convention of the CPU checkers (no FMA
         * contraction, correctly rounded division and square root) — a pragma in front of the unmodified file and one
         * standard build option; the default is the vendor compiler's own latitude, like a stock run of the reference. */
        const char* strict = getenv("BMFR_OPENCL_STRICT_FP");
        const int strict_fp = strict && strict[0] == '1';
        const char* pragma = "#pragma OPENCL FP_CONTRACT OFF\n";
        const char* srcs[2] = {pragma, (const char*)kBmfrClSource};
        const size_t lens[2] = {strlen(pragma), kBmfrClSourceLen};
        if (strict_fp) strncat(opts, " -cl-fp32-correctly-rounded-divide-sqrt", sizeof(opts) - strlen(opts) - 1);
        s->prog = cl.CreateProgramWithSource(s->ctx, strict_fp ? 2 : 1, strict_fp ? srcs : srcs + 1, strict_fp ? lens : lens + 1, &err);
        CL_TRY(err, "clCreateProgramWithSource");
        err = cl.BuildProgram(s->prog, 1, &s->dev, opts, NULL, NULL);
        if (err != CL_SUCCESS) {
            size_t n = 0;
            int off = snprintf(g_error, sizeof(g_error), "clBuildProgram failed (%d): ", err);
            cl.GetProgramBuildInfo(s->prog, s->dev, CL_PROGRAM_BUILD_LOG, sizeof(g_error) - 1 - (size_t)off, g_error + off, &n);
            goto fail;
        }
    }
    for (int i = 0; i < K_COUNT; ++i) {
        s->k[i] = cl.CreateKernel(s->prog, kKernelNames[i], &err);
        CL_TRY(err, kKernelNames[i]);
    }
    {   /* buffers, bmfr.cpp:315-343 */
        const size_t out = (size_t)s->Ww * s->Hw, m = (size_t)s->Wm * s->Hm, img = (size_t)s->W * s->H;
        const size_t tmp_elem = p->tmp_half ? 2 : 4;
        dbuf* dbs[] = {&s->normals, &s->positions, &s->noisy, &s->out, &s->result, &s->spp};
        const size_t dbytes[] = {out * 12, out * 12, out * 12, m * 12, out * 12, out};
        for (int i = 0; i < 6; ++i) {
            dbs[i]->a = make_buffer(s, CL_MEM_READ_WRITE, dbytes[i], &err); CL_TRY(err, "clCreateBuffer");
            dbs[i]->b = make_buffer(s, CL_MEM_READ_WRITE, dbytes[i], &err); CL_TRY(err, "clCreateBuffer");
        }
        s->in_buffer = make_buffer(s, CL_MEM_READ_WRITE, m * BMFR_CL_BUFFER_COUNT * tmp_elem, &err); CL_TRY(err, "clCreateBuffer");
        s->filtered = make_buffer(s, CL_MEM_READ_WRITE, out * 12, &err); CL_TRY(err, "clCreateBuffer");
        s->prev_pixels = make_buffer(s, CL_MEM_READ_WRITE, out * 8, &err); CL_TRY(err, "clCreateBuffer");
        s->accept = make_buffer(s, CL_MEM_READ_WRITE, out, &err); CL_TRY(err, "clCreateBuffer");
        s->albedo = make_buffer(s, CL_MEM_READ_ONLY, img * 12, &err); CL_TRY(err, "clCreateBuffer");
        s->tone_mapped = make_buffer(s, CL_MEM_READ_WRITE, img * 12, &err); CL_TRY(err, "clCreateBuffer");
        s->weights = make_buffer(s, CL_MEM_READ_WRITE, (size_t)s->NB * (BMFR_CL_BUFFER_COUNT - 3) * 3 * 4, &err); CL_TRY(err, "clCreateBuffer");
        s->mins_maxs = make_buffer(s, CL_MEM_READ_WRITE, (size_t)s->NB * 6 * 8, &err); CL_TRY(err, "clCreateBuffer");
    }
    {   /* static kernel arguments, bmfr.cpp:349-383 */
        const int bc = BMFR_CL_BUFFER_COUNT;
        const size_t r_size = (BMFR_CL_COMPRESSED_R ? (size_t)((bc - 2) * (bc - 1) / 2) : (size_t)((bc - 2) * (bc - 2))) * 16;  /* cl_float3 = 16 B */
        CL_TRY(cl.SetKernelArg(s->k[K_ACCUM_NOISY], 0, sizeof(cl_mem), &s->prev_pixels), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_ACCUM_NOISY], 1, sizeof(cl_mem), &s->accept), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_FITTER], 0, BMFR_CL_LOCAL_SIZE * sizeof(float), NULL), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_FITTER], 1, 1024 * sizeof(float), NULL), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_FITTER], 2, r_size, NULL), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_FITTER], 3, sizeof(cl_mem), &s->weights), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_FITTER], 4, sizeof(cl_mem), &s->mins_maxs), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_WEIGHTED_SUM], 0, sizeof(cl_mem), &s->weights), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_WEIGHTED_SUM], 1, sizeof(cl_mem), &s->mins_maxs), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_WEIGHTED_SUM], 2, sizeof(cl_mem), &s->filtered), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_ACCUM_FILTERED], 0, sizeof(cl_mem), &s->filtered), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_ACCUM_FILTERED], 1, sizeof(cl_mem), &s->prev_pixels), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_ACCUM_FILTERED], 2, sizeof(cl_mem), &s->accept), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_ACCUM_FILTERED], 3, sizeof(cl_mem), &s->albedo), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_ACCUM_FILTERED], 4, sizeof(cl_mem), &s->tone_mapped), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_TAA], 0, sizeof(cl_mem), &s->prev_pixels), "arg");
        CL_TRY(cl.SetKernelArg(s->k[K_TAA], 1, sizeof(cl_mem), &s->tone_mapped), "arg");
    }
    CL_TRY(cl.Finish(s->q), "clFinish");
    return s;
fail:
    oracle_destroy(s);
    return NULL;
}

static double event_ms(cl_event e, cl_ulong* start, cl_ulong* end) {
    cl_ulong t0 = 0, t1 = 0;
    cl.GetEventProfilingInfo(e, CL_PROFILING_COMMAND_START, sizeof(t0), &t0, NULL);
    cl.GetEventProfilingInfo(e, CL_PROFILING_COMMAND_END, sizeof(t1), &t1, NULL);
    if (start) *start = t0;
    if (end) *end = t1;
    return (double)(t1 - t0) * 1e-6;
}

/* One iteration of bmfr.cpp:417-485.  The uploads are blocking like the reference's and, like there, outside
 * every timer: the stage times are the kernels' own profiling events. */
static int run_frame(oracle_state* s, int frame, const float* albedo, const float* normals, const float* positions,
                     const float* noisy, const float cam_prev[16], const float pixel_offset[2]) {
    const size_t n = (size_t)s->W * s->H * 12;
    cl_event ev[K_COUNT] = {0};
    cl_mem a, b;
    float cam[16] = {0};
    if (cam_prev) memcpy(cam, cam_prev, sizeof(cam));
    a = db_cur(&s->normals);
    CL_TRY(cl.EnqueueWriteBuffer(s->q, s->albedo, 1, 0, n, albedo, 0, NULL, NULL), "write albedo");  /* bmfr.cpp:420-427 */
    CL_TRY(cl.EnqueueWriteBuffer(s->q, a, 1, 0, n, normals, 0, NULL, NULL), "write normals");
    a = db_cur(&s->positions);
    CL_TRY(cl.EnqueueWriteBuffer(s->q, a, 1, 0, n, positions, 0, NULL, NULL), "write positions");
    a = db_cur(&s->noisy);
    CL_TRY(cl.EnqueueWriteBuffer(s->q, a, 1, 0, n, noisy, 0, NULL, NULL), "write noisy");

    /* accumulate_noisy_data, args 2..13: bmfr.cpp:430-445 */
    {
        cl_kernel k = s->k[K_ACCUM_NOISY];
        cl_uint i = 2;
        a = db_cur(&s->normals); b = db_prev(&s->normals);
        CL_TRY(cl.SetKernelArg(k, i++, sizeof(cl_mem), &a), "arg"); CL_TRY(cl.SetKernelArg(k, i++, sizeof(cl_mem), &b), "arg");
        a = db_cur(&s->positions); b = db_prev(&s->positions);
        CL_TRY(cl.SetKernelArg(k, i++, sizeof(cl_mem), &a), "arg"); CL_TRY(cl.SetKernelArg(k, i++, sizeof(cl_mem), &b), "arg");
        a = db_cur(&s->noisy); b = db_prev(&s->noisy);
        CL_TRY(cl.SetKernelArg(k, i++, sizeof(cl_mem), &a), "arg"); CL_TRY(cl.SetKernelArg(k, i++, sizeof(cl_mem), &b), "arg");
        a = db_prev(&s->spp); b = db_cur(&s->spp);
        CL_TRY(cl.SetKernelArg(k, i++, sizeof(cl_mem), &a), "arg"); CL_TRY(cl.SetKernelArg(k, i++, sizeof(cl_mem), &b), "arg");
        CL_TRY(cl.SetKernelArg(k, i++, sizeof(cl_mem), &s->in_buffer), "arg");
        CL_TRY(cl.SetKernelArg(k, i++, 64, cam), "arg cam");
        CL_TRY(cl.SetKernelArg(k, i++, 8, pixel_offset), "arg offset");
        CL_TRY(cl.SetKernelArg(k, i++, sizeof(cl_int), &frame), "arg frame");
        const size_t g[2] = {(size_t)s->Wm, (size_t)s->Hm}, l[2] = {BMFR_CL_LOCAL_WIDTH, BMFR_CL_LOCAL_HEIGHT};  /* bmfr.cpp:245,247 */
        CL_TRY(cl.EnqueueNDRangeKernel(s->q, k, 2, NULL, g, l, 0, NULL, &ev[K_ACCUM_NOISY]), "accumulate_noisy_data");
    }
    {   /* fitter, args 5..6: bmfr.cpp:449-451 */
        cl_kernel k = s->k[K_FITTER];
        CL_TRY(cl.SetKernelArg(k, 5, sizeof(cl_mem), &s->in_buffer), "arg");
        CL_TRY(cl.SetKernelArg(k, 6, sizeof(cl_int), &frame), "arg");
        const size_t g[1] = {(size_t)BMFR_CL_LOCAL_SIZE * s->NB}, l[1] = {BMFR_CL_LOCAL_SIZE};  /* bmfr.cpp:248-249 */
        CL_TRY(cl.EnqueueNDRangeKernel(s->q, k, 1, NULL, g, l, 0, NULL, &ev[K_FITTER]), "fitter");
    }
    const size_t og[2] = {(size_t)s->Ww, (size_t)s->Hw}, ol[2] = {BMFR_CL_LOCAL_WIDTH, BMFR_CL_LOCAL_HEIGHT};  /* bmfr.cpp:246-247 */
    {   /* weighted_sum, args 3..6: bmfr.cpp:455-459 */
        cl_kernel k = s->k[K_WEIGHTED_SUM];
        a = db_cur(&s->normals); CL_TRY(cl.SetKernelArg(k, 3, sizeof(cl_mem), &a), "arg");
        a = db_cur(&s->positions); CL_TRY(cl.SetKernelArg(k, 4, sizeof(cl_mem), &a), "arg");
        a = db_cur(&s->noisy); CL_TRY(cl.SetKernelArg(k, 5, sizeof(cl_mem), &a), "arg");
        CL_TRY(cl.SetKernelArg(k, 6, sizeof(cl_int), &frame), "arg");
        CL_TRY(cl.EnqueueNDRangeKernel(s->q, k, 2, NULL, og, ol, 0, NULL, &ev[K_WEIGHTED_SUM]), "weighted_sum");
    }
    {   /* accumulate_filtered_data, args 5..8: bmfr.cpp:463-467 */
        cl_kernel k = s->k[K_ACCUM_FILTERED];
        a = db_cur(&s->spp); CL_TRY(cl.SetKernelArg(k, 5, sizeof(cl_mem), &a), "arg");
        a = db_prev(&s->out); CL_TRY(cl.SetKernelArg(k, 6, sizeof(cl_mem), &a), "arg");
        a = db_cur(&s->out); CL_TRY(cl.SetKernelArg(k, 7, sizeof(cl_mem), &a), "arg");
        CL_TRY(cl.SetKernelArg(k, 8, sizeof(cl_int), &frame), "arg");
        CL_TRY(cl.EnqueueNDRangeKernel(s->q, k, 2, NULL, og, ol, 0, NULL, &ev[K_ACCUM_FILTERED]), "accumulate_filtered_data");
    }
    {   /* taa, args 2..4: bmfr.cpp:471-474 */
        cl_kernel k = s->k[K_TAA];
        a = db_cur(&s->result); CL_TRY(cl.SetKernelArg(k, 2, sizeof(cl_mem), &a), "arg");
        a = db_prev(&s->result); CL_TRY(cl.SetKernelArg(k, 3, sizeof(cl_mem), &a), "arg");
        CL_TRY(cl.SetKernelArg(k, 4, sizeof(cl_int), &frame), "arg");
        CL_TRY(cl.EnqueueNDRangeKernel(s->q, k, 2, NULL, og, ol, 0, NULL, &ev[K_TAA]), "taa");
    }
    CL_TRY(cl.Finish(s->q), "clFinish");
    {   /* profiling, bmfr.cpp:488-506: per-kernel event durations; total = K1 start .. taa end */
        cl_ulong t_start = 0, t_end = 0;
        for (int i = 0; i < K_COUNT; ++i) s->ms[i] = event_ms(ev[i], i == 0 ? &t_start : NULL, i == K_COUNT - 1 ? &t_end : NULL);
        s->ms[5] = (double)(t_end - t_start) * 1e-6;
        for (int i = 0; i < K_COUNT; ++i) cl.ReleaseEvent(ev[i]);
    }
    /* swap all double buffers, bmfr.cpp:483-484 */
    s->normals.swapped ^= 1; s->positions.swapped ^= 1; s->noisy.swapped ^= 1; s->out.swapped ^= 1; s->result.swapped ^= 1;
    s->spp.swapped ^= 1;
    return 0;
fail:
    for (int i = 0; i < K_COUNT; ++i)
        if (ev[i]) cl.ReleaseEvent(ev[i]);
    return -1;
}

int oracle_frame(oracle_state* s, int frame, const float* albedo, const float* normals, const float* positions,
                 const float* noisy, const float cam_prev[16], const float pixel_offset[2]) {
    if (!s || !albedo || !normals || !positions || !noisy || !pixel_offset || (frame > 0 && !cam_prev)) return -1;
    return run_frame(s, frame, albedo, normals, positions, noisy, cam_prev, pixel_offset);
}

const void* oracle_buffer(oracle_state* s, int id, size_t* bytes) {
    if (!s || id < 0 || id >= 12) return NULL;
    const size_t npix = (size_t)s->W * s->H;
    cl_mem m = NULL;
    size_t n = 0;
    switch (id) {  /* after the swap the buffers the last frame wrote are the "previous" halves */
        case ORACLE_BUF_NOISY_ACC: m = db_prev(&s->noisy); n = npix * 12; break;
        case ORACLE_BUF_SPP: m = db_prev(&s->spp); n = npix; break;
        case ORACLE_BUF_PREV_PIXELS: m = s->prev_pixels; n = npix * 8; break;
        case ORACLE_BUF_ACCEPT: m = s->accept; n = npix; break;
        case ORACLE_BUF_WEIGHTS: m = s->weights; n = (size_t)s->NB * (BMFR_CL_BUFFER_COUNT - 3) * 3 * 4; break;
        case ORACLE_BUF_MINS_MAXS: m = s->mins_maxs; n = (size_t)s->NB * BMFR_CL_FEATURES_SCALED * 8; break;
        case ORACLE_BUF_FILTERED: m = s->filtered; n = npix * 12; break;
        case ORACLE_BUF_ACCUM: m = db_prev(&s->out); n = npix * 12; break;
        case ORACLE_BUF_TONE_MAPPED: m = s->tone_mapped; n = npix * 12; break;
        case ORACLE_BUF_RESULT: m = db_prev(&s->result); n = npix * 12; break;
        default: break;  /* tmp_data is destroyed by the fitter; the noise tile is never materialised on the device */
    }
    if (!m) return NULL;
    if (s->host_bytes[id] < n) {
        free(s->host[id]);
        s->host[id] = malloc(n);
        s->host_bytes[id] = s->host[id] ? n : 0;
    }
    if (!s->host[id]) return NULL;
    if (cl.EnqueueReadBuffer(s->q, m, 1, 0, n, s->host[id], 0, NULL, NULL) != CL_SUCCESS) return NULL;
    if (bytes) *bytes = n;
    return s->host[id];
}

void oracle_stage_ms(oracle_state* s, double ms[6]) {
    for (int i = 0; i < 6; ++i) ms[i] = s ? s->ms[i] : 0.0;
}
