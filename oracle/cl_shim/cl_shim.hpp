// TEST INFRASTRUCTURE — not part of the product.
//
// A minimal OpenCL-C-as-C++ shim: just enough of the OpenCL C language (vector types with the
// swizzles used, address-space qualifiers, built-ins, work-item functions, work-group barriers) to
// compile the reference's own kernel file /root/reference/opencl/bmfr.cl with g++ and run it on CPU
// cores.  The kernel source is NOT copied into this repository: oracle/cl_shim/build_ref.py reads
// it from /root/reference at build time and writes only into oracle/_ref/.
//
// Semantics fixed here where OpenCL leaves them open (same as oracle/bmfr_oracle.c):
//   dot() left to right; IEEE +,-,*,/ and sqrt (build with -ffp-contract=off); powr = powf;
//   convert_int_rtn saturates; fmin/fmax are C fminf/fmaxf.
#pragma once
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

namespace clshim {

// ---------------------------------------------------------------- work-item context
struct WorkItem {
    int global_id[2];
    int local_id[2];
    int group_id[2];
};
extern thread_local WorkItem g_wi;
void wg_barrier();  // fiber yield; implemented in cl_host.cpp

// run-time values behind the -D macros of bmfr.cpp:205-232
struct Config {
    int image_width, image_height, workset_width, workset_height, margin_width, margin_height;
    double noise_amount;
    float blend_alpha, second_blend_alpha, taa_blend_alpha;
    double position_limit_squared, normal_limit_squared;
};
extern Config g_cfg;

// ---------------------------------------------------------------- vector types
struct int2 {
    int x, y;
    int2() = default;
    int2(int a, int b) : x(a), y(b) {}
};
struct float2 {
    float x, y;
    float2() = default;
    float2(float a) : x(a), y(a) {}
    float2(float a, float b) : x(a), y(b) {}
};
struct float3 {
    float x, y, z, pad_;  // 16 bytes like cl_float3 (bmfr.cpp:356-358)
    float3() = default;
    float3(float a) : x(a), y(a), z(a), pad_(0.f) {}
    float3(float a, float b, float c) : x(a), y(b), z(c), pad_(0.f) {}
};
struct int3 {
    int x, y, z;
};

struct float4;
// `.xyz` of a float4 as an assignable member that shares storage with x,y,z
struct xyz_proxy {
    float v[3];
    operator float3() const { return float3(v[0], v[1], v[2]); }
    xyz_proxy& operator=(const float3& f) {
        v[0] = f.x; v[1] = f.y; v[2] = f.z;
        return *this;
    }
};
struct float4 {
    union {
        struct { float x, y, z, w; };
        xyz_proxy xyz;
    };
    float4() = default;
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};
// strided 4-element picks of a float16 (.s048c etc.), read-only
template <int A, int B, int C, int D>
struct pick4 {
    operator float4() const {
        const float* s = reinterpret_cast<const float*>(this);
        return float4(s[A], s[B], s[C], s[D]);
    }
};
struct float16 {
    union {
        float s[16];
        pick4<0, 4, 8, 12> s048c;
        pick4<1, 5, 9, 13> s159d;
        pick4<2, 6, 10, 14> s26ae;
        pick4<3, 7, 11, 15> s37bf;
    };
};
struct half {
    uint16_t bits;
};
typedef unsigned char uchar;

// ---------------------------------------------------------------- operators
#define CLSHIM_VEC2_OPS(T, S)                                                              \
    static inline T operator+(T a, T b) { return T(a.x + b.x, a.y + b.y); }                \
    static inline T operator-(T a, T b) { return T(a.x - b.x, a.y - b.y); }                \
    static inline T operator*(T a, T b) { return T(a.x * b.x, a.y * b.y); }                \
    static inline T operator/(T a, T b) { return T(a.x / b.x, a.y / b.y); }                \
    static inline T operator+(T a, S b) { return T(a.x + b, a.y + b); }                    \
    static inline T operator-(T a, S b) { return T(a.x - b, a.y - b); }                    \
    static inline T operator-(S a, T b) { return T(a - b.x, a - b.y); }                    \
    static inline T operator*(T a, S b) { return T(a.x * b, a.y * b); }                    \
    static inline T operator*(S a, T b) { return T(a * b.x, a * b.y); }                    \
    static inline T operator/(T a, S b) { return T(a.x / b, a.y / b); }                    \
    static inline T& operator+=(T& a, T b) { a = a + b; return a; }                        \
    static inline T& operator-=(T& a, T b) { a = a - b; return a; }                        \
    static inline T& operator+=(T& a, S b) { a = a + b; return a; }                        \
    static inline T& operator/=(T& a, S b) { a = a / b; return a; }
CLSHIM_VEC2_OPS(float2, float)
CLSHIM_VEC2_OPS(int2, int)

static inline float3 operator+(float3 a, float3 b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline float3 operator-(float3 a, float3 b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float3 operator*(float3 a, float3 b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float3 operator/(float3 a, float3 b) { return float3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline float3 operator*(float a, float3 b) { return float3(a * b.x, a * b.y, a * b.z); }
static inline float3 operator*(float3 a, float b) { return float3(a.x * b, a.y * b, a.z * b); }
static inline float3 operator/(float3 a, float b) { return float3(a.x / b, a.y / b, a.z / b); }
static inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
static inline float3& operator/=(float3& a, float b) { a = a / b; return a; }
// component-wise relational result, used only by select_lt below
static inline int3 operator<(float3 a, float b) { return int3{a.x < b ? -1 : 0, a.y < b ? -1 : 0, a.z < b ? -1 : 0}; }
// `c < 0.f ? 0.f : c` on a float3 (bmfr.cl:750) — the one vector ternary of the file; C++ cannot
// overload ?: so build_ref.py rewrites that expression to this call.
static inline float3 vec_ternary(int3 m, float a, float3 b) { return float3(m.x ? a : b.x, m.y ? a : b.y, m.z ? a : b.z); }

// ---------------------------------------------------------------- built-ins
static inline float dot(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float dot(float4 a, float4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
static inline float fmin(float a, float b) { return ::fminf(a, b); }
static inline float fmax(float a, float b) { return ::fmaxf(a, b); }
static inline float3 fmin(float3 a, float3 b) { return float3(::fminf(a.x, b.x), ::fminf(a.y, b.y), ::fminf(a.z, b.z)); }
static inline float3 fmax(float3 a, float3 b) { return float3(::fmaxf(a.x, b.x), ::fmaxf(a.y, b.y), ::fmaxf(a.z, b.z)); }
// OpenCL max(x,y) = y if x < y else x ; min(x,y) = y if y < x else x
static inline float max(float a, float b) { return a < b ? b : a; }
static inline float min(float a, float b) { return b < a ? b : a; }
static inline int max(int a, int b) { return a < b ? b : a; }
static inline int min(int a, int b) { return b < a ? b : a; }
static inline float3 max(float a, float3 b) { return float3(max(a, b.x), max(a, b.y), max(a, b.z)); }
static inline float sqrt(float a) { return ::sqrtf(a); }
static inline float fabs(float a) { return ::fabsf(a); }
static inline int abs(int a) { return a < 0 ? -a : a; }
static inline int isnan(float a) { return a != a; }
static inline float3 powr(float3 a, float b) { return float3(::powf(a.x, b), ::powf(a.y, b), ::powf(a.z, b)); }
static inline float clamp1(float v, float lo, float hi) { return ::fminf(::fmaxf(v, lo), hi); }
static inline float3 clamp(float3 v, float lo, float hi) { return float3(clamp1(v.x, lo, hi), clamp1(v.y, lo, hi), clamp1(v.z, lo, hi)); }
static inline float3 clamp(float3 v, float3 lo, float3 hi) { return float3(clamp1(v.x, lo.x, hi.x), clamp1(v.y, lo.y, hi.y), clamp1(v.z, lo.z, hi.z)); }

static inline float convert_float(unsigned int a) { return (float)a; }
static inline float convert_float(int a) { return (float)a; }
static inline float convert_float(uchar a) { return (float)a; }
static inline float convert_float(double a) { return (float)a; }
static inline float convert_float(float a) { return a; }
static inline float2 convert_float2(int2 a) { return float2((float)a.x, (float)a.y); }
static inline int floor_sat(float v) {
    if (v != v) return 0;
    const float f = ::floorf(v);
    if (f >= 2147483648.f) return INT_MAX;
    if (f <= -2147483648.f) return INT_MIN;
    return (int)f;
}
static inline int2 convert_int2_rtn(float2 a) { return int2(floor_sat(a.x), floor_sat(a.y)); }
static inline uchar convert_uchar_sat_rte(float v) {
    if (v != v) return 0;
    long r = ::lrintf(v);
    return (uchar)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

uint16_t f32_to_f16_rte(float f);
float f16_to_f32(uint16_t h);
static inline float vload_half(size_t i, const half* p) { return f16_to_f32(p[i].bits); }
static inline void vstore_half(float v, size_t i, half* p) { p[i].bits = f32_to_f16_rte(v); }

static inline int get_global_id(int d) { return g_wi.global_id[d]; }
static inline int get_local_id(int d) { return g_wi.local_id[d]; }
static inline int get_group_id(int d) { return g_wi.group_id[d]; }
#define CLK_LOCAL_MEM_FENCE 1
#define CLK_GLOBAL_MEM_FENCE 2
static inline void barrier(int) { wg_barrier(); }

}  // namespace clshim

// ---------------------------------------------------------------- qualifiers / keywords
#define __kernel
#define __global
#define __constant const
#define restrict __restrict__
// Work-group-local VARIABLE declarations (bmfr.cl:503,659) become per-OS-thread statics: one OS
// thread runs one work-group at a time, its 256 work-items being fibers of that thread.
// build_ref.py rewrites those two declarations to CLSHIM_WG_LOCAL; as a pointer qualifier
// (__local float* p) the keyword is dropped.
#define CLSHIM_WG_LOCAL static thread_local
#define __local
