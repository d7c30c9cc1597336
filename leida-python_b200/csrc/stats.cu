// Post-hoc statistics of the LEiDA pipeline on the GPU: the per-state two-group tests the reference
// runs on the occupancy and dwell-time tables after the hot path (stats/_stats.py:54-269, called from
// _leida.py:336-342) -- moments + Levene statistic of every variable, and the permutation null distributions
// of the independent-samples t statistic (pooled or Welch) and of the paired mean difference, batched over
// variables x permutations.
//
// One thread = one permutation of one variable.  The variable's observations sit in shared memory (centred by
// the pooled mean: both statistics are shift invariant and the sums stay small), every thread walks them once:
//   independent test: a uniformly random split is drawn by selection sampling (Knuth's Algorithm S: element i
//     joins group 1 with probability (n1 - taken) / (n - i)), so no permutation is ever materialised; group 1's
//     sum and sum of squares give both groups' means and variances;
//   paired test: every pair keeps or swaps its two values with probability 1/2.
// Random numbers: Philox4x32-10 keyed by the caller's seed, counter = (permutation, variable, position / 4).
// For exact tests (all splits enumerated by the host) and for bit-comparable tests against the CPU oracle the
// caller passes the splits explicitly instead.
#include "common.cuh"
#include <math.h>

namespace leida {

constexpr int STATS_MAX_N = 4096;        // pooled observations per variable held in shared memory
constexpr int STATS_THREADS = 128;

// ---- Philox4x32-10 (Salmon et al., SC'11) ---------------------------------------------------------------
struct Philox {
    uint32_t c[4], k[2];
    __device__ __forceinline__ void round_() {
        const uint32_t lo0 = 0xD2511F53u * c[0], hi0 = __umulhi(0xD2511F53u, c[0]);
        const uint32_t lo1 = 0xCD9E8D57u * c[2], hi1 = __umulhi(0xCD9E8D57u, c[2]);
        const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    __device__ __forceinline__ void generate(uint64_t seed, uint32_t a, uint32_t b, uint32_t d) {
        c[0] = a; c[1] = b; c[2] = d; c[3] = 0u;
        k[0] = (uint32_t)seed; k[1] = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) round_();
    }
};

// fixed-order block sum (deterministic): every thread returns the total
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += scratch[w];
    return t;
}

// t statistic of a split from group 1's sum and sum of squares (data centred, T = total sum, Q = total squares)
__device__ __forceinline__ double t_from_sums(double s1, double q1, double T, double Q, int n1, int n2, bool equal_var) {
    const double s2 = T - s1, q2 = Q - q1;
    const double m1 = s1 / n1, m2 = s2 / n2;
    const double ssd1 = q1 - s1 * m1, ssd2 = q2 - s2 * m2;
    double denom;
    if (equal_var) denom = sqrt((ssd1 + ssd2) / (double)(n1 + n2 - 2) * (1.0 / n1 + 1.0 / n2));
    else denom = sqrt(ssd1 / (double)(n1 - 1) / n1 + ssd2 / (double)(n2 - 1) / n2);
    return (m1 - m2) / denom;
}

// out[var][8] = {mean1, mean2, sum of squared deviations 1, 2, Levene W (center = mean), n1, n2, 0}
__global__ void __launch_bounds__(STATS_THREADS)
k_stats_moments(const double* __restrict__ values, int n_vars, const int32_t* __restrict__ idx1, int n1,
                const int32_t* __restrict__ idx2, int n2, double* __restrict__ out) {
    __shared__ double scratch[STATS_THREADS / 32];
    const int var = blockIdx.x;
    const int32_t* idx[2] = {idx1, idx2};
    const int nn[2] = {n1, n2};
    double mean[2], ssd[2], zbar[2], zssd[2];
    for (int g = 0; g < 2; ++g) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nn[g]; i += blockDim.x) s += values[(long long)idx[g][i] * n_vars + var];
        mean[g] = block_sum(s, scratch) / nn[g];
        double d2 = 0.0, z = 0.0;
        for (int i = threadIdx.x; i < nn[g]; i += blockDim.x) {
            const double d = values[(long long)idx[g][i] * n_vars + var] - mean[g];
            d2 += d * d;
            z += fabs(d);
        }
        ssd[g] = block_sum(d2, scratch);
        zbar[g] = block_sum(z, scratch) / nn[g];
        double zz = 0.0;
        for (int i = threadIdx.x; i < nn[g]; i += blockDim.x) {
            const double d = fabs(values[(long long)idx[g][i] * n_vars + var] - mean[g]) - zbar[g];
            zz += d * d;
        }
        zssd[g] = block_sum(zz, scratch);
    }
    if (threadIdx.x == 0) {
        const double ntot = (double)n1 + n2;
        const double zall = (zbar[0] * n1 + zbar[1] * n2) / ntot;
        const double numer = (ntot - 2.0) * (n1 * (zbar[0] - zall) * (zbar[0] - zall) + n2 * (zbar[1] - zall) * (zbar[1] - zall));
        double* o = out + (long long)var * 8;
        o[0] = mean[0]; o[1] = mean[1]; o[2] = ssd[0]; o[3] = ssd[1];
        o[4] = numer / (zssd[0] + zssd[1]);
        o[5] = n1; o[6] = n2; o[7] = 0.0;
    }
}

// grid = (ceil(n_perm / STATS_THREADS), n_vars).  counts[var][3] += {null <= t_obs, null >= t_obs, |null| >= |t_obs|}
__global__ void __launch_bounds__(STATS_THREADS)
k_perm_ttest_ind(const double* __restrict__ values, int n_vars, const int32_t* __restrict__ idx1, int n1,
                 const int32_t* __restrict__ idx2, int n2, const uint8_t* __restrict__ equal_var, int n_perm,
                 const int32_t* __restrict__ perms, unsigned long long seed, double* __restrict__ t_obs,
                 int32_t* __restrict__ counts, double* __restrict__ null_out) {
    __shared__ double x[STATS_MAX_N];
    __shared__ double scratch[STATS_THREADS / 32];
    const int var = blockIdx.y, n = n1 + n2;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = values[(long long)(i < n1 ? idx1[i] : idx2[i - n1]) * n_vars + var];
        x[i] = v;
        s += v;
    }
    const double centre = block_sum(s, scratch) / n;
    double q = 0.0, t = 0.0, s1o = 0.0, q1o = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = x[i] - centre;
        x[i] = v;
        t += v;
        q += v * v;
        if (i < n1) { s1o += v; q1o += v * v; }
    }
    const double T = block_sum(t, scratch), Q = block_sum(q, scratch);
    const double S1 = block_sum(s1o, scratch), Q1 = block_sum(q1o, scratch);
    const bool ev = equal_var[var] != 0;
    const double tobs = t_from_sums(S1, Q1, T, Q, n1, n2, ev);
    if (blockIdx.x == 0 && threadIdx.x == 0) t_obs[var] = tobs;
    __syncthreads();
    const int perm = blockIdx.x * blockDim.x + threadIdx.x;
    int c_le = 0, c_ge = 0, c_two = 0;
    if (perm < n_perm) {
        double s1 = 0.0, q1 = 0.0;
        if (perms) {
            const int32_t* pi = perms + (long long)perm * n;
            for (int i = 0; i < n1; ++i) {
                const double v = x[pi[i]];
                s1 += v;
                q1 += v * v;
            }
        } else {
            Philox rng;
            int taken = 0;
            for (int i = 0; i < n && taken < n1; i += 4) {
                rng.generate(seed, (uint32_t)perm, (uint32_t)var, (uint32_t)(i >> 2));
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = i + u;
                    if (e < n && taken < n1) {
                        // P(take) = (n1 - taken) / (n - e), decided in integers: r uniform on [0, 2^32)
                        if ((unsigned long long)rng.c[u] * (unsigned)(n - e) < ((unsigned long long)(n1 - taken) << 32)) {
                            const double v = x[e];
                            s1 += v;
                            q1 += v * v;
                            ++taken;
                        }
                    }
                }
            }
        }
        const double tn = t_from_sums(s1, q1, T, Q, n1, n2, ev);
        if (null_out) null_out[(long long)var * n_perm + perm] = tn;
        c_le = tn <= tobs;
        c_ge = tn >= tobs;
        c_two = (tn <= -fabs(tobs)) || (tn >= fabs(tobs));
    }
    c_le = warp_sum_i(c_le); c_ge = warp_sum_i(c_ge); c_two = warp_sum_i(c_two);
    if ((threadIdx.x & 31) == 0) {
        if (c_le) atomicAdd(&counts[var * 3], c_le);
        if (c_ge) atomicAdd(&counts[var * 3 + 1], c_ge);
        if (c_two) atomicAdd(&counts[var * 3 + 2], c_two);
    }
}

// Paired test: statistic = mean(x1) - mean(x2) = mean(d); a permutation flips the sign of d_i where pair i is swapped.
// counts[var][2] += {null <= obs + gamma, null >= obs - gamma}, gamma = 100 eps |obs| (scipy.stats.permutation_test).
__global__ void __launch_bounds__(STATS_THREADS)
k_perm_paired(const double* __restrict__ values, int n_vars, const int32_t* __restrict__ idx1,
              const int32_t* __restrict__ idx2, int n, int n_perm, const uint8_t* __restrict__ swaps,
              unsigned long long seed, double* __restrict__ obs_out, int32_t* __restrict__ counts,
              double* __restrict__ null_out) {
    __shared__ double d[STATS_MAX_N];
    __shared__ double scratch[STATS_THREADS / 32];
    const int var = blockIdx.y;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = values[(long long)idx1[i] * n_vars + var] - values[(long long)idx2[i] * n_vars + var];
        d[i] = v;
        s += v;
    }
    const double obs = block_sum(s, scratch) / n;
    if (blockIdx.x == 0 && threadIdx.x == 0) obs_out[var] = obs;
    __syncthreads();
    const double gamma = fabs(2.220446049250313e-16 * 100.0 * obs);
    const int perm = blockIdx.x * blockDim.x + threadIdx.x;
    int c_le = 0, c_ge = 0;
    if (perm < n_perm) {
        double acc = 0.0;
        if (swaps) {
            const uint8_t* sw = swaps + (long long)perm * n;
            for (int i = 0; i < n; ++i) acc += sw[i] ? -d[i] : d[i];
        } else {
            Philox rng;
            for (int i = 0; i < n; i += 128) {
                rng.generate(seed, (uint32_t)perm, (uint32_t)var, (uint32_t)(i >> 7));
                for (int u = 0; u < 128 && i + u < n; ++u) acc += ((rng.c[u >> 5] >> (u & 31)) & 1u) ? -d[i + u] : d[i + u];
            }
        }
        const double v = acc / n;
        if (null_out) null_out[(long long)var * n_perm + perm] = v;
        c_le = v <= obs + gamma;
        c_ge = v >= obs - gamma;
    }
    c_le = warp_sum_i(c_le); c_ge = warp_sum_i(c_ge);
    if ((threadIdx.x & 31) == 0) {
        if (c_le) atomicAdd(&counts[var * 2], c_le);
        if (c_ge) atomicAdd(&counts[var * 2 + 1], c_ge);
    }
}

}  // namespace leida

using namespace leida;

extern "C" int leida_stats_moments_f64(const double* values, int64_t n_rows, int n_vars, const int32_t* idx1, int n1,
                                       const int32_t* idx2, int n2, double* out, leida_stream_t stream) {
    LEIDA_REQUIRE(values && idx1 && idx2 && out, "null pointer");
    LEIDA_REQUIRE(n_rows >= 1 && n_vars >= 1 && n_vars <= 65535 && n1 >= 1 && n2 >= 1, "sizes");
    k_stats_moments<<<n_vars, STATS_THREADS, 0, (cudaStream_t)stream>>>(values, n_vars, idx1, n1, idx2, n2, out);
    return check_launch("k_stats_moments");
}

extern "C" int leida_perm_ttest_ind_f64(const double* values, int64_t n_rows, int n_vars, const int32_t* idx1, int n1,
                                        const int32_t* idx2, int n2, const uint8_t* equal_var, int n_perm,
                                        const int32_t* perms, uint64_t seed, double* t_obs, int32_t* counts,
                                        double* null_out, leida_stream_t stream) {
    LEIDA_REQUIRE(values && idx1 && idx2 && equal_var && t_obs && counts, "null pointer");
    LEIDA_REQUIRE(n_rows >= 1 && n_vars >= 1 && n_vars <= 65535 && n1 >= 2 && n2 >= 2 && n1 + n2 <= STATS_MAX_N && n_perm >= 1,
                  "sizes (2 <= n1, n2; n1 + n2 <= 4096)");
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 3 * (size_t)n_vars, st));
    dim3 grid((unsigned)((n_perm + STATS_THREADS - 1) / STATS_THREADS), (unsigned)n_vars);
    k_perm_ttest_ind<<<grid, STATS_THREADS, 0, st>>>(values, n_vars, idx1, n1, idx2, n2, equal_var, n_perm, perms,
                                                     (unsigned long long)seed, t_obs, counts, null_out);
    return check_launch("k_perm_ttest_ind");
}

extern "C" int leida_perm_paired_f64(const double* values, int64_t n_rows, int n_vars, const int32_t* idx1,
                                     const int32_t* idx2, int n, int n_perm, const uint8_t* swaps, uint64_t seed,
                                     double* obs, int32_t* counts, double* null_out, leida_stream_t stream) {
    LEIDA_REQUIRE(values && idx1 && idx2 && obs && counts, "null pointer");
    LEIDA_REQUIRE(n_rows >= 1 && n_vars >= 1 && n_vars <= 65535 && n >= 1 && n <= STATS_MAX_N && n_perm >= 1, "sizes (n <= 4096)");
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 2 * (size_t)n_vars, st));
    dim3 grid((unsigned)((n_perm + STATS_THREADS - 1) / STATS_THREADS), (unsigned)n_vars);
    k_perm_paired<<<grid, STATS_THREADS, 0, st>>>(values, n_vars, idx1, idx2, n, n_perm, swaps, (unsigned long long)seed, obs,
                                                  counts, null_out);
    return check_launch("k_perm_paired");
}
