// STAGED mode with params.reference_order = 1: the fitter and the weighted sum evaluated in the
// reference's own operation order, so that EVERY buffer of the frame loop is bit-identical to the
// reference kernels' arithmetic (as pinned by the oracle), including the reference's shipped default
// USE_HALF_PRECISION_IN_TMP_DATA = 1 (bmfr.cpp:88), where every store into tmp_data rounds to fp16
// (bmfr.cl:255-258, 472, 540, 651).
//
// This is the compatibility path, not the fast one: it keeps the reference's schedule — one 256-thread
// group per block, thread `id` owns rows id + 256 s (IN_ACCESS, bmfr.cl:90-97), 13 column iterations,
// the fixed 256 -> 64 -> 8 -> 1 reduction trees (bmfr.cl:26-87) and a block barrier wherever the
// reference has one — because the summation order IS the result.  The matrix still lives in registers /
// shared memory per column instead of making 193 passes over global memory.
// Compiled with --fmad=false: every a*b+c below is two correctly rounded operations, left to right,
// like the oracle (gcc -ffp-contract=off).
#include <cuda_fp16.h>

#include "bmfr_kernels.h"

#include "bmfr_device.cuh"

#define RO_THREADS 256  // LOCAL_SIZE, bmfr.cpp:116
#define RO_SUB 4        // BLOCK_PIXELS / LOCAL_SIZE
#define RO_R_EDGE 11    // buffers - 2, bmfr.cpp:221
#define RO_R_SIZE 66    // packed upper triangle of 11 x 11, bmfr.cl:100-103

template <bool HALF>
__device__ __forceinline__ float tmp_load(const void* t, size_t i) {
    return HALF ? __half2float(static_cast<const __half*>(t)[i]) : static_cast<const float*>(t)[i];
}
template <bool HALF>
__device__ __forceinline__ void tmp_store(void* t, size_t i, float v) {
    if (HALF) static_cast<__half*>(t)[i] = __float2half_rn(v);  // vstore_half: round to nearest even
    else static_cast<float*>(t)[i] = v;
}

// R_ACCESS with COMPRESSED_R = 1, bmfr.cl:100-103
__device__ __forceinline__ int r_access(int x, int y) { return (RO_R_SIZE - (RO_R_EDGE - y) * (RO_R_EDGE - y + 1) / 2) + x - y; }

// scale(), bmfr.cl:200-205 — with the division the reference writes
__device__ __forceinline__ float scale_exact(float value, float mn, float mx) {
    if (fabsf(mx - mn) > 1.0f) return (value - mn) / (mx - mn);
    return value - mn;
}

// parallel_reduction_{sum,min,max}, bmfr.cl:26-87.  v[id] holds this thread's term on entry.
template <int OP>  // 0 sum, 1 min, 2 max
__device__ __forceinline__ float ro_combine(float a, float b) { return OP == 0 ? a + b : OP == 1 ? fminf(a, b) : fmaxf(a, b); }
template <int OP>
__device__ __forceinline__ float ro_reduce(float* v, int id) {
    __syncthreads();
    if (id < 64) {
        if (OP == 0) v[id] += v[id + 64] + v[id + 128] + v[id + 192];
        else v[id] = ro_combine<OP>(ro_combine<OP>(ro_combine<OP>(v[id], v[id + 64]), v[id + 128]), v[id + 192]);
    }
    __syncthreads();
    if (id < 8) {
        if (OP == 0) v[id] += v[id + 8] + v[id + 16] + v[id + 24] + v[id + 32] + v[id + 40] + v[id + 48] + v[id + 56];
        else {
            float r = v[id];
#pragma unroll
            for (int k = 8; k < 64; k += 8) r = ro_combine<OP>(r, v[id + k]);
            v[id] = r;
        }
    }
    __syncthreads();
    float r;
    if (OP == 0) r = v[0] + v[1] + v[2] + v[3] + v[4] + v[5] + v[6] + v[7];
    else {
        r = v[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) r = ro_combine<OP>(r, v[k]);
    }
    __syncthreads();  // v is reused by the caller
    return r;
}

// --------------------------------------------------------------------------------------------
// fitter, bmfr.cl:490-700, in the reference's order.  One CTA per block of the margin grid.
// --------------------------------------------------------------------------------------------
template <bool HALF>
__global__ void __launch_bounds__(RO_THREADS) k2_fitter_reference_order_kernel(const __grid_constant__ KParams P) {
    __shared__ float sum_vec[RO_THREADS];
    __shared__ float u_vec[BMFR_BLOCK_PIXELS];
    __shared__ float r_mat[RO_R_SIZE][3];
    __shared__ float s_vec_length, s_u_length_squared;
    __shared__ float s_divider[3];
    const int id = threadIdx.x;
    const int bx = blockIdx.x, by = P.by0 + blockIdx.y;
    const int group = by * P.blocks_x + bx;
    constexpr int buffers = BMFR_BUFFER_COUNT;
    const size_t base = (size_t)((by - P.by0) * P.blocks_x + bx) * buffers * BMFR_BLOCK_PIXELS;
    void* tmp = P.tmp_data;
#define RO_IN(fb, sv) (base + (size_t)(fb) * BMFR_BLOCK_PIXELS + (sv) * RO_THREADS + id)

    for (int i = id; i < RO_R_SIZE * 3; i += RO_THREADS) (&r_mat[0][0])[i] = 0.f;

    // (i) min / max scaling of the six scaled features, bmfr.cl:511-542
    for (int fb = BMFR_FEATURES_NOT_SCALED; fb < buffers - 3; ++fb) {
        float tmp_max = -CUDART_INF_F, tmp_min = CUDART_INF_F;
        float vals[RO_SUB];
#pragma unroll
        for (int sv = 0; sv < RO_SUB; ++sv) {
            vals[sv] = tmp_load<HALF>(tmp, RO_IN(fb, sv));
            tmp_max = fmaxf(vals[sv], tmp_max);
            tmp_min = fminf(vals[sv], tmp_min);
        }
        sum_vec[id] = tmp_max;
        const float block_max = ro_reduce<2>(sum_vec, id);
        sum_vec[id] = tmp_min;
        const float block_min = ro_reduce<1>(sum_vec, id);
        if (id == 0) {
            const int index = (group * BMFR_FEATURES_SCALED + fb - BMFR_FEATURES_NOT_SCALED) * 2;
            P.mins_maxs[index + 0] = block_min;
            P.mins_maxs[index + 1] = block_max;
            P.mins_inv[index + 0] = block_min;
            P.mins_inv[index + 1] = scale_factor(block_min, block_max);
        }
#pragma unroll
        for (int sv = 0; sv < RO_SUB; ++sv) tmp_store<HALF>(tmp, RO_IN(fb, sv), scale_exact(vals[sv], block_min, block_max));
    }

    // (ii) Householder columns, bmfr.cl:546-656
    for (int col = 0; col < buffers; ++col) {
        const int col_limited = col < buffers - 3 ? col : buffers - 3;
        float tmp_sum_value = 0.f;
#pragma unroll
        for (int sv = 0; sv < RO_SUB; ++sv) {
            const float t = tmp_load<HALF>(tmp, RO_IN(col, sv));
            const int index = id + sv * RO_THREADS;
            u_vec[index] = t;
            if (index >= col_limited + 1) tmp_sum_value += t * t;
        }
        sum_vec[id] = tmp_sum_value;
        float vec_length = ro_reduce<0>(sum_vec, id);  // also makes u_vec visible

        // bmfr.cl:574-600
        float r_value;
        if (id < col) {
            r_value = u_vec[id];
        } else if (id == col) {
            float u_length_squared = vec_length;
            vec_length = sqrtf(vec_length + u_vec[col_limited] * u_vec[col_limited]);
            u_vec[col_limited] -= vec_length;
            u_length_squared += u_vec[col_limited] * u_vec[col_limited];
            s_vec_length = vec_length;
            s_u_length_squared = u_length_squared;
            r_value = vec_length;
        } else {
            r_value = 0.0f;
        }
        {
            const int id_limited = id < buffers - 3 ? id : buffers - 3;
            float* r = r_mat[r_access(col_limited, id_limited)];
            // threads 10.. all address R(col_limited, 10): the reference has the same benign collision
            // (SURVEY H2c); only R(10,10) receives conflicting values and nothing reads it
            if (col < buffers - 3) {
                if (id <= buffers - 3) r[0] = r[1] = r[2] = r_value;  // store_r_mat_broadcast
            } else if (id < buffers - 3) {
                r[col - buffers + 3] = r_value;                        // store_r_mat_channel, rows 0..9 of the colour column
            }
        }
        __syncthreads();
        const float u_length_squared = s_u_length_squared;

        for (int fb = col_limited + 1; fb < buffers; ++fb) {  // bmfr.cl:606-655
            float cache[RO_SUB];
            tmp_sum_value = 0.f;
#pragma unroll
            for (int sv = 0; sv < RO_SUB; ++sv) {
                const int index = id + sv * RO_THREADS;
                cache[sv] = 0.f;
                if (index >= col_limited) {
                    float t = tmp_load<HALF>(tmp, RO_IN(fb, sv));
                    if (col == 0 && fb < buffers - 3)  // add_random, bmfr.cl:623-627 — an fp64 expression rounded once
                        t = (float)((double)t + P.noise[(fb - 1) * BMFR_BLOCK_PIXELS + index]);
                    cache[sv] = t;
                    tmp_sum_value += t * u_vec[index];
                }
            }
            sum_vec[id] = tmp_sum_value;
            const float dot = ro_reduce<0>(sum_vec, id);
#pragma unroll
            for (int sv = 0; sv < RO_SUB; ++sv) {
                const int index = id + sv * RO_THREADS;
                if (index >= col_limited) {
                    float store_value = cache[sv];
                    store_value -= 2 * u_vec[index] * dot / u_length_squared;  // bmfr.cl:650
                    tmp_store<HALF>(tmp, RO_IN(fb, sv), store_value);
                }
            }
        }
        __syncthreads();  // u_vec is rewritten by the next column
    }

    // (iii) back substitution, bmfr.cl:659-692
    for (int i = RO_R_EDGE - 2; i >= 0; --i) {
        if (id == 0) {
            const float* d = r_mat[r_access(i, i)];
            s_divider[0] = d[0]; s_divider[1] = d[1]; s_divider[2] = d[2];
        }
        __syncthreads();
        if (id < RO_R_EDGE && id >= i) {
            float* v = r_mat[r_access(id, i)];
            v[0] = v[0] / s_divider[0]; v[1] = v[1] / s_divider[1]; v[2] = v[2] / s_divider[2];
        }
        __syncthreads();
        if (id == 0)
            for (int j = i + 1; j < RO_R_EDGE - 1; ++j) {
                float* v = r_mat[r_access(RO_R_EDGE - 1, i)];
                const float* v2 = r_mat[r_access(j, i)];
                v[0] = v[0] - v2[0]; v[1] = v[1] - v2[1]; v[2] = v[2] - v2[2];
            }
        __syncthreads();
        if (id < RO_R_EDGE && i >= id) {
            float* v = r_mat[r_access(i, id)];
            const float* v2 = r_mat[r_access(RO_R_EDGE - 1, i)];
            const float t0 = v[0] * v2[0], t1 = v[1] * v2[1], t2 = v[2] * v2[2];
            v[0] = t0; v[1] = t1; v[2] = t2;
        }
        __syncthreads();
    }
    if (id < buffers - 3) {  // bmfr.cl:694-699
        const float* w = r_mat[r_access(RO_R_EDGE - 1, id)];
        float* out = P.weights + ((size_t)group * (buffers - 3) + id) * 3;
        out[0] = w[0]; out[1] = w[1]; out[2] = w[2];
    }
#undef RO_IN
}

// --------------------------------------------------------------------------------------------
// weighted_sum, bmfr.cl:703-758, in the reference's order (division in scale(), no contraction).
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k3_weighted_sum_reference_order_kernel(const __grid_constant__ KParams P) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = P.py0 + blockIdx.y * 8 + threadIdx.y;
    if (x >= P.W || y >= P.py1) return;
    const unsigned int lp = pix_index(P, x, y);
    const int g = k3_group(P, x, y);
    const f3 n = load_f3(P.cur_normals, lp), p = load_f3(P.cur_positions, lp);
    const float features[BMFR_FEATURES] = {1.f, n.x, n.y, n.z, p.x, p.y, p.z, p.x * p.x, p.y * p.y, p.z * p.z};
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int fb = 0; fb < BMFR_FEATURES; ++fb) {
        float feature = features[fb];
        if (fb >= BMFR_FEATURES_NOT_SCALED) {
            const float* mm = P.mins_maxs + ((size_t)g * BMFR_FEATURES_SCALED + fb - BMFR_FEATURES_NOT_SCALED) * 2;
            feature = scale_exact(feature, __ldg(mm), __ldg(mm + 1));
        }
        const float* w = P.weights + ((size_t)g * BMFR_FEATURES + fb) * 3;
        c0 += __ldg(w) * feature;
        c1 += __ldg(w + 1) * feature;
        c2 += __ldg(w + 2) * feature;
    }
    store_f3(P.filtered, lp, make_f3(c0 < 0.f ? 0.f : c0, c1 < 0.f ? 0.f : c1, c2 < 0.f ? 0.f : c2));  // bmfr.cl:750
}

cudaError_t launch_k2_reference_order(const KParams& P, bool half, cudaStream_t st) {
    const dim3 grid(P.blocks_x, P.by1 - P.by0);
    if (half) k2_fitter_reference_order_kernel<true><<<grid, RO_THREADS, 0, st>>>(P);
    else k2_fitter_reference_order_kernel<false><<<grid, RO_THREADS, 0, st>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_k3_reference_order(const KParams& P, cudaStream_t st) {
    k3_weighted_sum_reference_order_kernel<<<dim3((P.W + 31) / 32, (P.py1 - P.py0 + 7) / 8), dim3(32, 8), 0, st>>>(P);
    return cudaGetLastError();
}
