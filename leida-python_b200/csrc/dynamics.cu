// Stage 4 of the LEiDA hot path on sm_100a: integer counts behind the dynamics metrics.
// Replaces the per-subject loops of pyleida/dynamics_metrics/_dynamics_metrics.py:
//   occupancy counts (:131-134, value_counts), run counts (:447-453, itertools.groupby),
//   transition pair counts (:242-245).  The fp64 ratios are formed on the host from these exact
//   integers with the same operations as the reference, so they are bit-identical.
#include "common.cuh"

namespace leida {

// One CTA per (segment, label column).  Shared-memory histograms; int32 atomics are native in smem.
__global__ void __launch_bounds__(256)
k_dynamics_counts(const int32_t* __restrict__ labels, long long label_pitch, const int32_t* __restrict__ col_k,
                  const long long* __restrict__ seg_off, const long long* __restrict__ seg_rows, int n_segments,
                  int kmax, int32_t* __restrict__ occ, int32_t* __restrict__ runs, int32_t* __restrict__ trans) {
    extern __shared__ int sh[];
    int* h_occ = sh;                 // [kmax]
    int* h_run = sh + kmax;          // [kmax]
    int* h_tr = sh + 2 * kmax;       // [kmax*kmax]
    const int g = blockIdx.x, col = blockIdx.y;
    const int K = col_k[col];
    const int32_t* lab = labels + (long long)col * label_pitch;
    for (int i = threadIdx.x; i < 2 * kmax + kmax * kmax; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const long long b = seg_off[g], e = seg_off[g + 1];
    for (long long i = b + threadIdx.x; i < e; i += blockDim.x) {
        const int cur = lab[seg_rows ? seg_rows[i] : i];
        if (cur < 0 || cur >= K) continue;   // out-of-range labels are ignored (never produced by the path)
        atomicAdd(&h_occ[cur], 1);
        int prev = -1;
        if (i > b) prev = lab[seg_rows ? seg_rows[i - 1] : i - 1];
        if (prev != cur) atomicAdd(&h_run[cur], 1);
        if (i + 1 < e) {
            const int nxt = lab[seg_rows ? seg_rows[i + 1] : i + 1];
            if (nxt >= 0 && nxt < K) atomicAdd(&h_tr[cur * kmax + nxt], 1);
        }
    }
    __syncthreads();
    const long long base = (long long)col * n_segments + g;
    for (int i = threadIdx.x; i < kmax; i += blockDim.x) {
        occ[base * kmax + i] = h_occ[i];
        runs[base * kmax + i] = h_run[i];
    }
    for (int i = threadIdx.x; i < kmax * kmax; i += blockDim.x) trans[base * kmax * kmax + i] = h_tr[i];
}

}  // namespace leida

using namespace leida;

extern "C" int leida_dynamics_counts_i32(const int32_t* labels, int64_t label_pitch, int n_cols, const int32_t* col_k,
                                         const int64_t* seg_off, const int64_t* seg_rows, int n_segments, int kmax,
                                         int32_t* occ, int32_t* runs, int32_t* trans, leida_stream_t stream) {
    LEIDA_REQUIRE(labels && col_k && seg_off && occ && runs && trans, "null pointer");
    LEIDA_REQUIRE(n_cols >= 1 && n_cols <= 65535, "1 <= n_cols <= 65535");
    LEIDA_REQUIRE(n_segments >= 0, "n_segments >= 0");
    LEIDA_REQUIRE(kmax >= 1 && kmax <= 100, "1 <= kmax <= 100");
    if (n_segments == 0) return LEIDA_OK;
    const size_t smem = (size_t)(2 * kmax + kmax * kmax) * sizeof(int);
    dim3 grid((unsigned)n_segments, (unsigned)n_cols);
    k_dynamics_counts<<<grid, 256, smem, (cudaStream_t)stream>>>(labels, label_pitch, col_k, (const long long*)seg_off,
                                                                  (const long long*)seg_rows, n_segments, kmax, occ, runs,
                                                                  trans);
    return check_launch("k_dynamics_counts");
}
