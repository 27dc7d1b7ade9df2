// synth-v1: deterministic synthetic path-tracer output for the BMFR hot path.
//
// The reference reads its inputs from a 19 GB dataset (.exr frames plus camera_matrices.h,
// /root/reference/opencl/bmfr.cpp:44-53,145-163) that is not in the repository.  This header
// defines the stand-in scene once, as code shared by a host build (g++ -ffp-contract=off) and a
// device build (nvcc --fmad=false).  Only +,-,*,/ and sqrt are used on floats, all correctly
// rounded on both sides, so the two twins produce bit-identical frames.
//
// Per frame it produces what bmfr.cpp:420-427,440-444 hands to the kernels:
//   albedo, shading normal, world position, 1-spp noisy (demodulated) colour : W*H*3 fp32 each,
//   camera matrix of the frame (row-vector convention clip_j = sum_i p_i*M[i][j],
//   bmfr.cl:343-347), pixel offset (bmfr.cl:352-355).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SYNTH_HD __host__ __device__ __forceinline__
#else
#define SYNTH_HD static inline
#endif

#define SYNTH_SEED_DEFAULT 0x424D4652u  // "BMFR"

// Camera of one frame, floats only (the double-precision trig lives in synth_camera()).
struct SynthCamera {
    float o[3];      // eye
    float r[3];      // right
    float d[3];      // down (image y grows downwards, matching uv.y*H in bmfr.cl:352)
    float f[3];      // forward
    float thx, thy;  // tan(half fov) in x and y
};

struct SynthPixel {
    float albedo[3];
    float normal[3];
    float position[3];
    float color[3];
};

SYNTH_HD uint32_t synth_hash(uint32_t a) {
    a ^= a >> 16; a *= 0x7feb352du;
    a ^= a >> 15; a *= 0x846ca68bu;
    a ^= a >> 16;
    return a;
}

SYNTH_HD float synth_sqrt(float x) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return __builtin_sqrtf(x);
#endif
}

SYNTH_HD float synth_floor(float x) {
#if defined(__CUDA_ARCH__)
    return floorf(x);
#else
    return __builtin_floorf(x);
#endif
}

SYNTH_HD float synth_dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

#define SYNTH_NUM_SPHERES 5
// x, y, z, radius, albedo rgb
SYNTH_HD void synth_sphere(int i, float* c, float* rad, float* alb) {
    const float S[SYNTH_NUM_SPHERES][7] = {
        {-3.0f, 1.2f, 4.0f, 1.2f, 0.85f, 0.30f, 0.25f},
        {0.5f, 0.8f, 1.5f, 0.8f, 0.30f, 0.80f, 0.35f},
        {3.2f, 1.6f, 6.0f, 1.6f, 0.25f, 0.40f, 0.88f},
        {-0.8f, 0.5f, -1.5f, 0.5f, 0.90f, 0.85f, 0.30f},
        {5.0f, 0.9f, 0.5f, 0.9f, 0.70f, 0.70f, 0.70f},
    };
    c[0] = S[i][0]; c[1] = S[i][1]; c[2] = S[i][2];
    *rad = S[i][3];
    alb[0] = S[i][4]; alb[1] = S[i][5]; alb[2] = S[i][6];
}

// nearest positive hit of ray o + t*dir with sphere i, or -1
SYNTH_HD float synth_hit_sphere(const float* o, const float* dir, const float* c, float rad) {
    float oc[3] = {o[0] - c[0], o[1] - c[1], o[2] - c[2]};
    float a = synth_dot3(dir, dir);
    float b = synth_dot3(oc, dir);
    float cc = synth_dot3(oc, oc) - rad * rad;
    float disc = b * b - a * cc;
    if (disc <= 0.f) return -1.f;
    float t = (-b - synth_sqrt(disc)) / a;
    return t > 1e-3f ? t : -1.f;
}

// One primary ray.  x,y pixel; frame for the 1-spp noise.
SYNTH_HD SynthPixel synth_pixel(const SynthCamera& cam, int W, int H, int x, int y, int frame, uint32_t seed) {
    SynthPixel px;
    const float u = ((float)x + 0.5f) / (float)W * 2.f - 1.f;
    const float v = ((float)y + 0.5f) / (float)H * 2.f - 1.f;
    const float su = u * cam.thx, sv = v * cam.thy;
    float dir[3];
    for (int k = 0; k < 3; ++k) dir[k] = (cam.f[k] + su * cam.r[k]) + sv * cam.d[k];

    float tbest = 1e30f;
    int what = -1;  // -1 sky, 0 floor, 1 back wall, 2 left wall, 3 right wall, 10+i sphere
    if (dir[1] < 0.f) {
        float t = -cam.o[1] / dir[1];
        if (t > 0.f && t < tbest) { tbest = t; what = 0; }
    }
    if (dir[2] > 0.f) {
        float t = (14.f - cam.o[2]) / dir[2];
        if (t > 0.f && t < tbest) { tbest = t; what = 1; }
    }
    if (dir[0] < 0.f) {
        float t = (-10.f - cam.o[0]) / dir[0];
        if (t > 0.f && t < tbest) { tbest = t; what = 2; }
    }
    if (dir[0] > 0.f) {
        float t = (10.f - cam.o[0]) / dir[0];
        if (t > 0.f && t < tbest) { tbest = t; what = 3; }
    }
    float sc[3], srad, salb[3];
    for (int i = 0; i < SYNTH_NUM_SPHERES; ++i) {
        float c[3], rad, alb[3];
        synth_sphere(i, c, &rad, alb);
        float t = synth_hit_sphere(cam.o, dir, c, rad);
        if (t > 0.f && t < tbest) {
            tbest = t; what = 10 + i;
            sc[0] = c[0]; sc[1] = c[1]; sc[2] = c[2]; srad = rad;
            salb[0] = alb[0]; salb[1] = alb[1]; salb[2] = alb[2];
        }
    }
    // walls are 9 high; above them (or too far away) is sky
    float p[3] = {cam.o[0] + tbest * dir[0], cam.o[1] + tbest * dir[1], cam.o[2] + tbest * dir[2]};
    if (what >= 1 && what <= 3 && p[1] > 9.f) what = -1;
    if (what == 0 && tbest > 60.f) what = -1;

    const uint32_t h = synth_hash(synth_hash(synth_hash((uint32_t)x + 0x9E3779B9u * (uint32_t)y) ^
                                             (0x85EBCA6Bu * (uint32_t)(frame + 1))) ^ seed);
    const float noise = 1.f + 0.8f * ((float)(h >> 8) * (1.f / 16777216.f) - 0.5f);

    if (what < 0) {
        const float ts = 80.f;
        for (int k = 0; k < 3; ++k) { px.position[k] = cam.o[k] + ts * dir[k]; px.normal[k] = 0.f; }
        px.albedo[0] = 0.55f; px.albedo[1] = 0.70f; px.albedo[2] = 0.90f;
        const float g = 0.9f - 0.3f * (v * 0.5f + 0.5f);
        px.color[0] = g * noise; px.color[1] = g * noise; px.color[2] = g * noise;
        return px;
    }

    float n[3], alb[3];
    if (what == 0) {
        n[0] = 0.f; n[1] = 1.f; n[2] = 0.f;
        const int chk = ((int)synth_floor(p[0]) + (int)synth_floor(p[2])) & 1;
        const float a = chk ? 0.85f : 0.25f;
        alb[0] = a; alb[1] = a * 0.95f; alb[2] = a * 0.85f;
    } else if (what == 1) {
        n[0] = 0.f; n[1] = 0.f; n[2] = -1.f;
        const int st = ((int)synth_floor(p[0] * 0.5f)) & 1;
        alb[0] = st ? 0.80f : 0.45f; alb[1] = st ? 0.55f : 0.50f; alb[2] = st ? 0.35f : 0.75f;
    } else if (what == 2) {
        n[0] = 1.f; n[1] = 0.f; n[2] = 0.f;
        const int st = ((int)synth_floor(p[1]) + (int)synth_floor(p[2] * 0.5f)) & 1;
        alb[0] = st ? 0.75f : 0.35f; alb[1] = st ? 0.30f : 0.65f; alb[2] = st ? 0.30f : 0.40f;
    } else if (what == 3) {
        n[0] = -1.f; n[1] = 0.f; n[2] = 0.f;
        const int st = ((int)synth_floor(p[1] * 2.f)) & 1;
        alb[0] = st ? 0.30f : 0.60f; alb[1] = st ? 0.45f : 0.60f; alb[2] = st ? 0.80f : 0.35f;
    } else {
        const float inv = 1.f / srad;
        n[0] = (p[0] - sc[0]) * inv; n[1] = (p[1] - sc[1]) * inv; n[2] = (p[2] - sc[2]) * inv;
        alb[0] = salb[0]; alb[1] = salb[1]; alb[2] = salb[2];
    }

    // direct light from one point light with hard sphere shadows + constant ambient
    const float L[3] = {2.5f, 8.5f, -3.f};
    float l[3] = {L[0] - p[0], L[1] - p[1], L[2] - p[2]};
    const float d2 = synth_dot3(l, l);
    const float dl = synth_sqrt(d2);
    float ndl = synth_dot3(n, l) / dl;
    if (ndl < 0.f) ndl = 0.f;
    float so[3] = {p[0] + 1e-2f * n[0], p[1] + 1e-2f * n[1], p[2] + 1e-2f * n[2]};
    float lit = 1.f;
    for (int i = 0; i < SYNTH_NUM_SPHERES; ++i) {
        float c[3], rad, a3[3];
        synth_sphere(i, c, &rad, a3);
        float t = synth_hit_sphere(so, l, c, rad);
        if (t > 0.f && t < 1.f) lit = 0.f;
    }
    const float e = 0.30f + lit * (55.f * ndl / d2);
    for (int k = 0; k < 3; ++k) {
        px.albedo[k] = alb[k];
        px.normal[k] = n[k];
        px.position[k] = p[k];
    }
    px.color[0] = e * noise;
    px.color[1] = e * 0.97f * noise;
    px.color[2] = e * 0.90f * noise;
    return px;
}
