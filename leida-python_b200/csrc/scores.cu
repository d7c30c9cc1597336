// Clustering quality scores of identify_states (pyleida/clustering/_clustering.py:331-339, 934-973):
//   Dunn index (dunn_fast, cosine distances), silhouette_score(metric='cosine'), davies_bouldin_score.
// SURVEY section 8f-1 ("next" row).  All label columns (one per K of the sweep) are scored together.
//
//   * cosine distances follow sklearn.metrics.pairwise.cosine_distances on float32 input: rows are
//     normalised in float32, d = clip(1 - <xi, xj>, 0, 2) in float32, d(i,i) = 0.
//   * silhouette needs only the mean distance of a row to every cluster, and
//       sum_{j in c} d(i,j) = n_c - <x^_i, G_c>,   G_c = sum_{j in c} x^_j
//     so it is O(n * sum K * N), not O(n^2); accumulated in fp64 / exact integers.
//   * Dunn needs min/max over all pairs: one tiled n^2 pass shared by all K (upper triangle only).
//   * Davies-Bouldin (Euclidean): mean distance of the members to their centroid, fp64.
#include "common.cuh"
#include <math.h>

namespace leida {

constexpr double kScoreScale = 1099511627776.0;   // 2^40 fixed point for order-independent sums

// rows[sel[i]] normalised like sklearn.preprocessing.normalize on float32 -> xs (n, NP), pad 0.
__global__ void k_scores_normalize(const float* __restrict__ X, long long XP, int N, const long long* __restrict__ sel,
                                   long long n, float* __restrict__ xs) {
    const int lane = threadIdx.x & 31;
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const float* xr = X + (sel ? sel[i] : i) * XP;
    float ss = 0.f;
    for (int j = lane; j < N; j += 32) ss = fmaf(xr[j], xr[j], ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float nrm = sqrtf(ss);
    const float inv = nrm > 0.f ? 1.f / nrm : 0.f;   // sklearn leaves all-zero rows untouched
    for (int j = lane; j < XP; j += 32) xs[i * XP + j] = j < N ? xr[j] * inv : 0.f;
}

// G[col0[c] + label_c(i)][:] += x^_i  (fixed point), cnt[...] += 1.   One warp per row, all columns.
__global__ void k_scores_cluster_sums(const float* __restrict__ xs, long long XP, int N, long long n,
                                      const int32_t* __restrict__ labels, long long label_pitch,
                                      const long long* __restrict__ sel, int n_cols, const int32_t* __restrict__ col0,
                                      long long* __restrict__ G, int32_t* __restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const long long src = sel ? sel[i] : i;
    for (int c = 0; c < n_cols; ++c) {
        const int row = col0[c] + labels[(long long)c * label_pitch + src];
        for (int j = lane; j < N; j += 32)
            atomicAdd(reinterpret_cast<unsigned long long*>(G + (long long)row * XP + j),
                      (unsigned long long)__double2ll_rn((double)xs[i * XP + j] * kScoreScale));
        if (lane == 0) atomicAdd(&cnt[row], 1);
    }
}

// silhouette sum per label column: one warp per row, the row in registers (fp64).
constexpr int SCORE_MAX_COLS_PER_LANE = 16;
__global__ void __launch_bounds__(256)
k_scores_silhouette(const float* __restrict__ xs, long long XP, int N, long long n, const int32_t* __restrict__ labels,
                    long long label_pitch, const long long* __restrict__ sel, int n_cols,
                    const int32_t* __restrict__ col0, const int32_t* __restrict__ col_k,
                    const long long* __restrict__ G, const int32_t* __restrict__ cnt, long long* __restrict__ sil_sum) {
    const int lane = threadIdx.x & 31;
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const long long src = sel ? sel[i] : i;
    const bool in_regs = N <= 32 * SCORE_MAX_COLS_PER_LANE;
    double xv[SCORE_MAX_COLS_PER_LANE];
    double self = 0.0;
#pragma unroll
    for (int u = 0; u < SCORE_MAX_COLS_PER_LANE; ++u) {
        const int j = lane + 32 * u;
        xv[u] = (in_regs && j < N) ? (double)xs[i * XP + j] : 0.0;
    }
    if (in_regs) {
#pragma unroll
        for (int u = 0; u < SCORE_MAX_COLS_PER_LANE; ++u) self += xv[u] * xv[u];
    } else {
        for (int j = lane; j < N; j += 32) self += (double)xs[i * XP + j] * (double)xs[i * XP + j];
    }
    self = warp_sum(self);
    for (int c = 0; c < n_cols; ++c) {
        const int K = col_k[c], r0 = col0[c];
        const int own = labels[(long long)c * label_pitch + src];
        double a = 0.0, b = INFINITY;
        for (int k = 0; k < K; ++k) {
            const long long* g = G + (long long)(r0 + k) * XP;
            double dot = 0.0;
            if (in_regs) {
#pragma unroll
                for (int u = 0; u < SCORE_MAX_COLS_PER_LANE; ++u) {
                    const int j = lane + 32 * u;
                    if (j < N) dot = fma(xv[u], (double)g[j], dot);
                }
            } else {
                for (int j = lane; j < N; j += 32) dot = fma((double)xs[i * XP + j], (double)g[j], dot);
            }
            dot = warp_sum(dot) * (1.0 / kScoreScale);
            const double nk = (double)cnt[r0 + k];
            if (k == own) a = nk > 1.0 ? ((nk - 1.0) - (dot - self)) / (nk - 1.0) : 0.0;
            else if (nk > 0.0) b = fmin(b, (nk - dot) / nk);
        }
        if (lane == 0) {
            double s = 0.0;
            if (cnt[r0 + own] > 1 && isfinite(b)) {
                const double m = fmax(a, b);
                s = m > 0.0 ? (b - a) / m : 0.0;          // sklearn: nan_to_num of 0/0
            }
            atomicAdd(reinterpret_cast<unsigned long long*>(sil_sum + c), (unsigned long long)__double2ll_rn(s * kScoreScale));
        }
    }
}

// Dunn: one CTA per 64x64 tile of the upper triangle; fp32 dots (4x4 per thread), then per label
// column the tile's min non-zero inter-cluster and max intra-cluster distance.
constexpr int DT = 64;
__global__ void __launch_bounds__(256)
k_scores_dunn(const float* __restrict__ xs, long long XP, long long n, const int32_t* __restrict__ labels,
              long long label_pitch, const long long* __restrict__ sel, int n_cols,
              unsigned int* __restrict__ min_inter, unsigned int* __restrict__ max_intra) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;
    __shared__ float sa[DT][33], sb[DT][33];
    __shared__ int la[DT], lb[DT];
    __shared__ unsigned int s_min, s_max;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const long long i0 = (long long)bi * DT, j0 = (long long)bj * DT;
    for (int c0 = 0; c0 < XP; c0 += 32) {
        for (int e = threadIdx.x; e < DT * 32; e += 256) {
            const int r = e >> 5, c = e & 31;
            sa[r][c] = (i0 + r < n) ? xs[(i0 + r) * XP + c0 + c] : 0.f;
            sb[r][c] = (j0 + r < n) ? xs[(j0 + r) * XP + c0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < 32; ++c) {
            float av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = sa[ty * 4 + a][c];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = sb[tx * 4 + b][c];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
    float d[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const long long gi = i0 + ty * 4 + a, gj = j0 + tx * 4 + b;
            float v = fminf(fmaxf(1.f - acc[a][b], 0.f), 2.f);
            if (gi == gj) v = 0.f;
            d[a][b] = (gi < n && gj < n) ? v : -1.f;          // -1 = outside the matrix
        }
    for (int c = 0; c < n_cols; ++c) {
        __syncthreads();
        if (threadIdx.x < DT) {
            const long long gi = i0 + threadIdx.x;
            la[threadIdx.x] = gi < n ? labels[(long long)c * label_pitch + (sel ? sel[gi] : gi)] : -1;
        } else if (threadIdx.x < 2 * DT) {
            const long long gj = j0 + threadIdx.x - DT;
            lb[threadIdx.x - DT] = gj < n ? labels[(long long)c * label_pitch + (sel ? sel[gj] : gj)] : -2;
        }
        if (threadIdx.x == 0) { s_min = 0x7f800000u; s_max = 0u; }
        __syncthreads();
        float mn = INFINITY, mx = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float v = d[a][b];
                if (v < 0.f) continue;
                if (la[ty * 4 + a] == lb[tx * 4 + b]) mx = fmaxf(mx, v);
                else if (v != 0.f) mn = fminf(mn, v);
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if ((threadIdx.x & 31) == 0) {          // non-negative floats order like their bit patterns
            atomicMin(&s_min, __float_as_uint(mn));
            atomicMax(&s_max, __float_as_uint(mx));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (s_min != 0x7f800000u) atomicMin(&min_inter[c], s_min);
            if (s_max != 0u) atomicMax(&max_intra[c], s_max);
        }
    }
}

// Davies-Bouldin: sum of Euclidean distances of every row to the centroid of its cluster (fp64),
// per (label column, cluster), exact integer accumulation.  One warp per row, all label columns.
__global__ void __launch_bounds__(256)
k_scores_db(const float* __restrict__ X, long long XP, int N, long long M, const int32_t* __restrict__ labels,
            long long label_pitch, int n_cols, const int32_t* __restrict__ col0, const double* __restrict__ centroids,
            long long* __restrict__ dist_sum) {
    const int lane = threadIdx.x & 31;
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= M) return;
    const float* xr = X + i * XP;
    for (int c = 0; c < n_cols; ++c) {
        const int row = col0[c] + labels[(long long)c * label_pitch + i];
        const double* ce = centroids + (long long)row * N;
        double s = 0.0;
        for (int j = lane; j < N; j += 32) {
            const double t = (double)xr[j] - ce[j];
            s = fma(t, t, s);
        }
        s = warp_sum(s);
        if (lane == 0)
            atomicAdd(reinterpret_cast<unsigned long long*>(dist_sum + row), (unsigned long long)__double2ll_rn(sqrt(s) * kScoreScale));
    }
}

}  // namespace leida

using namespace leida;

extern "C" int leida_scores_cosine(const float* X, int64_t x_pitch, int N, const int64_t* sel, int64_t n,
                                   const int32_t* labels, int64_t label_pitch, int n_cols, const int32_t* col0,
                                   const int32_t* col_k, int total_clusters, float* xs, int64_t* G, int32_t* cnt,
                                   int64_t* sil_sum, uint32_t* min_inter, uint32_t* max_intra, leida_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_REQUIRE(X && labels && col0 && col_k && xs && G && cnt && sil_sum && min_inter && max_intra, "null pointer");
    LEIDA_REQUIRE(n >= 1 && N >= 1 && x_pitch >= N && x_pitch % 32 == 0 && n_cols >= 1 && total_clusters >= 1, "sizes");
    LEIDA_REQUIRE(n <= 65535LL * DT, "at most 4.1 M rows in the pairwise pass");
    LEIDA_CUDA_TRY(cudaMemsetAsync(G, 0, (size_t)total_clusters * x_pitch * sizeof(int64_t), st));
    LEIDA_CUDA_TRY(cudaMemsetAsync(cnt, 0, (size_t)total_clusters * sizeof(int32_t), st));
    LEIDA_CUDA_TRY(cudaMemsetAsync(sil_sum, 0, (size_t)n_cols * sizeof(int64_t), st));
    LEIDA_CUDA_TRY(cudaMemsetAsync(min_inter, 0x7f, (size_t)n_cols * sizeof(uint32_t), st));   // 0x7f7f7f7f > any distance <= 2
    LEIDA_CUDA_TRY(cudaMemsetAsync(max_intra, 0, (size_t)n_cols * sizeof(uint32_t), st));
    const unsigned wblocks = (unsigned)((n * 32 + 255) / 256);
    k_scores_normalize<<<wblocks, 256, 0, st>>>(X, x_pitch, N, (const long long*)sel, n, xs);
    k_scores_cluster_sums<<<wblocks, 256, 0, st>>>(xs, x_pitch, N, n, labels, label_pitch, (const long long*)sel, n_cols, col0,
                                                   (long long*)G, cnt);
    k_scores_silhouette<<<wblocks, 256, 0, st>>>(xs, x_pitch, N, n, labels, label_pitch, (const long long*)sel, n_cols, col0,
                                                 col_k, (const long long*)G, cnt, (long long*)sil_sum);
    const unsigned nb = (unsigned)((n + DT - 1) / DT);
    k_scores_dunn<<<dim3(nb, nb), 256, 0, st>>>(xs, x_pitch, n, labels, label_pitch, (const long long*)sel, n_cols, min_inter,
                                                max_intra);
    return check_launch("leida_scores_cosine", 4);
}

extern "C" int leida_scores_db(const float* X, int64_t x_pitch, int N, int64_t M, const int32_t* labels,
                               int64_t label_pitch, int n_cols, const int32_t* col0, const double* centroids,
                               int total_clusters, int64_t* dist_sum, leida_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_REQUIRE(X && labels && col0 && centroids && dist_sum, "null pointer");
    LEIDA_REQUIRE(M >= 1 && N >= 1 && x_pitch >= N && n_cols >= 1 && total_clusters >= 1, "sizes");
    LEIDA_CUDA_TRY(cudaMemsetAsync(dist_sum, 0, (size_t)total_clusters * sizeof(int64_t), st));
    k_scores_db<<<(unsigned)((M * 32 + 255) / 256), 256, 0, st>>>(X, x_pitch, N, M, labels, label_pitch, n_cols, col0, centroids,
                                                                  (long long*)dist_sum);
    return check_launch("k_scores_db");
}
