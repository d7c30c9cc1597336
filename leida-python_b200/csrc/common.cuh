// Shared helpers for the LEiDA sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "leida_b200.h"

namespace leida {

void set_error(const char* fmt, ...);
int  sm_count();
int  max_smem_optin();

void count_launches(int n);

inline int check_launch(const char* what, int n_launched = 1) {
    count_launches(n_launched);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return LEIDA_ERR_CUDA;
    }
    return LEIDA_OK;
}

#define LEIDA_CUDA_TRY(expr)                                                        \
    do {                                                                            \
        cudaError_t e__ = (expr);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            leida::set_error("%s: %s", #expr, cudaGetErrorString(e__));             \
            return LEIDA_ERR_CUDA;                                                  \
        }                                                                           \
    } while (0)

#define LEIDA_REQUIRE(cond, msg)                                                    \
    do {                                                                            \
        if (!(cond)) {                                                              \
            leida::set_error("invalid argument: %s (%s)", msg, #cond);              \
            return LEIDA_ERR_INVALID;                                               \
        }                                                                           \
    } while (0)

// Sequential fp64 dot product in the reference's order (scipy cdist 'cosine'; no FMA contraction).
__device__ __forceinline__ double dot_seq(const float* __restrict__ a, const float* __restrict__ b, int n) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s = __dadd_rn(s, __dmul_rn((double)a[j], (double)b[j]));
    return s;
}

__device__ __forceinline__ double cosine_distance(double dot, double na, double nb) {
    double c = dot / (na * nb);
    if (fabs(c) > 1.0) c = copysign(1.0, c);
    return 1.0 - c;
}

// Width (in packed centroid columns) of the slot / TMEM read that holds a run of K clusters.
__host__ __device__ __forceinline__ int tc_read_width(int K) {
    return K <= 4 ? 4 : K <= 8 ? 8 : K <= 12 ? 12 : K <= 16 ? 16 : K <= 20 ? 20 : K <= 24 ? 24 : K <= 32 ? 32 : K <= 40 ? 40
         : K <= 48 ? 48 : 64;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace leida
