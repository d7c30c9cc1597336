// Stage 3 of the LEiDA hot path on sm_100a: cosine k-means, many runs (K, restart) in lock-step.
// Replaces pyleida/clustering/_clustering.py:82-153 (KMeansLeida.fit), :162-207, :209-232.
//
// Design (DESIGN.md section "k-means"):
//   * assign: one CTA = 128 threads x RPT rows, streams X in 128-byte column chunks through a
//     cp.async ring (XOR-swizzled so row-per-thread LDS.128 is conflict free), keeps K x RPT fp32
//     accumulators in registers, centroid chunks are broadcast from shared memory.
//   * parity: the fp32 score is only trusted when the top-2 margin exceeds a rigorous fp32 error
//     bound; other rows go to a work list and are re-evaluated in fp64 with the reference's
//     arithmetic order (scipy cdist 'cosine': sequential dot, dot/(|u||v|), clip, 1-cos, first min).
//   * centroid sums are int64 fixed point (2^40) and updated incrementally by the rows whose label
//     changed: exact, order independent, identical for any number of GPUs.
#include "common.cuh"
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include <type_traits>
#include <vector>

namespace leida {

constexpr int kStateWords = LEIDA_KM_STATE_WORDS;
constexpr double kFracScale = 1099511627776.0;  // 2^40

__device__ __forceinline__ long long quantize(float v) { return __double2ll_rn((double)v * kFracScale); }

__device__ __forceinline__ void atomic_add_ll(long long* p, long long v) {
    atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct AssignParams {
    const float* X;
    const double* xnorm;
    long long M;
    int N;
    int XP;  // row pitch of X and cent (floats), multiple of 32
    const int32_t* run_k;
    const int32_t* run_crow0;
    const float* cent;
    const double* cnorm;
    long long* acc;
    int32_t* labels;
    long long* state;
    int32_t* worklist;  // [0] = count, entries (run,row) from worklist+8
    long long worklist_cap;
    float tau0;
    int32_t* labels_out;  // recheck only: when set, write the exact label here and leave the move to the caller
    int n_runs;           // recheck only: runs to scan when the queue overflowed
    int metric;           // 0: cosine distance (the LEiDA path), 1: Euclidean (KMeansLeida(metric='euclidean'), CUDA-core flow)
};

// Exact (fp64) distance of scipy's cdist for the two metrics the reference accepts (clustering/_clustering.py:61-75):
// 'cosine' from the dot product and the norms, 'euclidean' = sqrt(sum (x - c)^2) from the squared difference.
__device__ __forceinline__ double metric_distance(int metric, double acc, double xn, double cn) {
    return metric ? sqrt(acc) : cosine_distance(acc, xn, cn);
}
__device__ __forceinline__ double seq_acc(int metric, const float* __restrict__ a, const float* __restrict__ b, int n) {
    if (!metric) return dot_seq(a, b, n);
    double s = 0.0;
    for (int j = 0; j < n; ++j) {
        const double d = __dsub_rn((double)a[j], (double)b[j]);
        s = __dadd_rn(s, __dmul_rn(d, d));
    }
    return s;
}

// Move one row's fixed-point contribution between clusters.  Called by `nthreads` cooperating
// threads (`t` = index of the caller among them).
__device__ __forceinline__ void move_row(const AssignParams& p, long long row, int crow0, int old_l, int new_l, int t,
                                         int nthreads) {
    const float* xr = p.X + row * p.XP;
    const int AP = p.XP + 8;
    long long* an = p.acc + (long long)(crow0 + new_l) * AP;
    long long* ao = p.acc + (long long)(crow0 + (old_l < 0 ? 0 : old_l)) * AP;
    for (int j = t; j < p.N; j += nthreads) {
        const long long q = quantize(xr[j]);
        atomic_add_ll(an + j, q);
        if (old_l >= 0) atomic_add_ll(ao + j, -q);
    }
    if (t == 0) {
        atomic_add_ll(an + p.XP, 1);
        if (old_l >= 0) atomic_add_ll(ao + p.XP, -1);
    }
}

template <int KP, int RPT>
__global__ void __launch_bounds__(128) k_assign(const AssignParams p) {
    constexpr int TR = 128 * RPT;
    constexpr int STAGES = 3;
    extern __shared__ __align__(128) unsigned char smraw[];
    float* sx = reinterpret_cast<float*>(smraw);            // [STAGES][TR][32]
    float* sc = sx + STAGES * TR * 32;                      // [STAGES][KP][32]
    float* s_cinv = sc + STAGES * KP * 32;                  // [KP]
    int* s_list = reinterpret_cast<int*>(s_cinv + KP);      // [TR] packed (local row | old<<16 | new<<24)
    __shared__ int s_nchg;

    const int run = blockIdx.y;
    if (p.state[(long long)run * kStateWords + LEIDA_KM_ST_ACTIVE] == 0) return;
    const int K = p.run_k[run];
    const int crow0 = p.run_crow0[run];
    const int tid = threadIdx.x;
    const long long row_base = (long long)blockIdx.x * TR;
    const int nch = p.XP / 32;
    const float* cbase = p.cent + (long long)crow0 * p.XP;

    __shared__ float s_cmax;
    if (tid == 0) {
        s_nchg = 0;
        double m = 0.0;
        if (p.metric)
            for (int k = 0; k < K; ++k) m = fmax(m, p.cnorm[crow0 + k]);
        s_cmax = (float)m * 1.0000002f;            // rounded up
    }
    if (tid < KP) {
        // cosine: score = dot / |c| (argmax); Euclidean: score = dot - |c|^2 / 2 (argmax == argmin of |x - c|^2)
        float ci = 0.f;
        if (tid < K) {
            const double cn = p.cnorm[crow0 + tid];
            ci = p.metric ? (float)(-0.5 * cn * cn) : (cn > 0.0 ? (float)(1.0 / cn) : 0.f);
        } else if (p.metric) {
            ci = -INFINITY;
        }
        s_cinv[tid] = ci;
    }

    auto issue = [&](int ch, int stage) {
        float* dx = sx + stage * TR * 32;
#pragma unroll
        for (int it = 0; it < RPT * 8; ++it) {
            const int i = tid + it * 128;
            const int r = i >> 3, c = i & 7;
            long long row = row_base + r;
            const bool ok = row < p.M;
            if (!ok) row = p.M - 1;
            cp_async16(dx + r * 32 + ((c ^ (r & 7)) << 2), p.X + row * p.XP + ch * 32 + c * 4, ok);
        }
        float* dc = sc + stage * KP * 32;
        for (int i = tid; i < K * 8; i += 128) {
            const int r = i >> 3, c = i & 7;
            cp_async16(dc + r * 32 + c * 4, cbase + (long long)r * p.XP + ch * 32 + c * 4, true);
        }
    };

    float acc[RPT][KP];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[i][k] = 0.f;

    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nch) issue(s, s);
        cp_async_commit();
    }
    const int swz = tid & 7;
    for (int ch = 0; ch < nch; ++ch) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = ch + STAGES - 1;
        if (nxt < nch) issue(nxt, nxt % STAGES);
        cp_async_commit();
        const float* bx = sx + (ch % STAGES) * TR * 32;
        const float* bc = sc + (ch % STAGES) * KP * 32;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
            float4 xv[RPT];
#pragma unroll
            for (int i = 0; i < RPT; ++i)
                xv[i] = *reinterpret_cast<const float4*>(bx + (tid + i * 128) * 32 + ((c4 ^ swz) << 2));
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                const float4 cv = *reinterpret_cast<const float4*>(bc + k * 32 + c4 * 4);
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    acc[i][k] = fmaf(xv[i].x, cv.x, acc[i][k]);
                    acc[i][k] = fmaf(xv[i].y, cv.y, acc[i][k]);
                    acc[i][k] = fmaf(xv[i].z, cv.z, acc[i][k]);
                    acc[i][k] = fmaf(xv[i].w, cv.w, acc[i][k]);
                }
            }
        }
    }
    cp_async_wait<0>();

    int32_t* lab = p.labels + (long long)run * p.M;
    int n_changed_local = 0;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int lr = tid + i * 128;
        const long long row = row_base + lr;
        if (row >= p.M) continue;
        float best = -INFINITY, second = -INFINITY;
        int bk = 0;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            if (k < K) {
                const float s = p.metric ? acc[i][k] + s_cinv[k] : acc[i][k] * s_cinv[k];
                if (s > best) { second = best; best = s; bk = k; }
                else if (s > second) second = s;
            }
        }
        const float xn = (float)p.xnorm[row];
        // fp32 error of a score: cosine <= tau0/2 * |x|; Euclidean <= tau0/2 * (|x| |c| + |c|^2 / 2) with |c| <= cmax
        const bool trusted = (best - second) > p.tau0 * (p.metric ? (xn * s_cmax + s_cmax * s_cmax) : xn);
        if (trusted) {
            const int old_l = lab[row];
            if (old_l != bk) {
                lab[row] = bk;
                const int slot = atomicAdd(&s_nchg, 1);
                s_list[slot] = lr | ((old_l & 0xff) << 16) | (bk << 24);
                ++n_changed_local;
            }
        } else {
            const int slot = atomicAdd(p.worklist, 1);
            if (slot < p.worklist_cap) {
                int4* e = reinterpret_cast<int4*>(p.worklist + 8) + 2 * (long long)slot;
                e[0] = make_int4(run, (int32_t)row, -1, 0);       // no tensor-core argmax to compare with
                e[1] = make_int4(0, 0, 0, 0);                      // empty candidate mask = all K clusters
            } else {
                // work list full: resolve here, serially (slow, correct)
                const float* xr = p.X + row * p.XP;
                const double xnd = p.xnorm[row];
                double bd = INFINITY;
                int bkk = 0;
                for (int k = 0; k < K; ++k) {
                    const double d = metric_distance(p.metric, seq_acc(p.metric, xr, cbase + (long long)k * p.XP, p.N), xnd,
                                                     p.cnorm[crow0 + k]);
                    if (d < bd) { bd = d; bkk = k; }
                }
                const int old_l = lab[row];
                if (old_l != bkk) {
                    lab[row] = bkk;
                    const int sl = atomicAdd(&s_nchg, 1);
                    s_list[sl] = lr | ((old_l & 0xff) << 16) | (bkk << 24);
                    ++n_changed_local;
                }
                atomic_add_ll(p.state + (long long)run * kStateWords + LEIDA_KM_ST_RECHECKED, 1);
            }
        }
    }
    __syncthreads();
    const int nchg = s_nchg;
    for (int e = 0; e < nchg; ++e) {
        const int v = s_list[e];
        const int lr = v & 0xffff;
        int old_l = (v >> 16) & 0xff;
        if (old_l == 0xff) old_l = -1;
        const int new_l = (v >> 24) & 0xff;
        move_row(p, row_base + lr, crow0, old_l, new_l, tid, 128);
    }
    if (tid == 0 && nchg > 0) atomic_add_ll(p.state + (long long)run * kStateWords + LEIDA_KM_ST_CHANGED, nchg);
}

// fp64 re-evaluation of the rows the fast pass could not decide.  One warp per queue entry
// {run, row, fast argmax, margin/threshold, candidate mask (64 bits), -, -}: the row is held in registers as
// doubles (lanes stride the columns), the dot product with every CANDIDATE centroid (the clusters whose fast score
// was within the threshold of the best; empty mask = all K) is an fp64 FMA chain per lane plus a fixed-order
// butterfly sum, then the reference's formula 1 - dot/(|x||c|) with clipping, first minimum wins.  (scipy's own
// build sums the dot product in two SIMD lanes, so no particular fp64 summation order is "the" reference; any fp64
// order agrees with it to ~1e-16 and decides identically unless two distances coincide to that level.)
// CPL = columns per lane held in registers (rows up to 32*CPL columns); CPL = 0: rows stay in memory.
template <int CPL>
__global__ void __launch_bounds__(128) k_recheck(const AssignParams p) {
    constexpr int RECHECK_MAX_COLS_PER_LANE = CPL > 0 ? CPL : 1;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const long long queued = p.worklist[0];
    const long long count = queued > p.worklist_cap ? p.worklist_cap : queued;
    constexpr bool in_regs = CPL > 0;
    auto resolve = [&](int run, long long row, unsigned long long mask, int fast_k, int ratio_bits) {
        const int K = p.run_k[run];
        const int crow0 = p.run_crow0[run];
        const unsigned long long all = K >= 64 ? ~0ull : ((1ull << K) - 1ull);
        mask &= all;
        if (mask == 0) mask = all;
        const float* xr = p.X + row * p.XP;
        const double xn = p.xnorm[row];
        double xv[RECHECK_MAX_COLS_PER_LANE];
        if (in_regs) {
#pragma unroll
            for (int i = 0; i < RECHECK_MAX_COLS_PER_LANE; ++i) {
                const int j = lane + 32 * i;
                xv[i] = j < p.N ? (double)xr[j] : 0.0;
            }
        }
        double bd = INFINITY;
        int bk = 0;
        bool any_nan = false;
        // candidates in increasing k (first minimum wins), U at a time: two when that covers the mask (the usual
        // case: best and second), else four
        auto batch = [&](auto tag) {
            constexpr int U = decltype(tag)::value;
            int ks[U];
            int n = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (mask) { ks[u] = __ffsll((long long)mask) - 1; mask &= mask - 1; n = u + 1; }
                else ks[u] = ks[u > 0 ? u - 1 : 0];
            }
            double s[U];
#pragma unroll
            for (int u = 0; u < U; ++u) s[u] = 0.0;
            if (in_regs) {
                float cv[U][RECHECK_MAX_COLS_PER_LANE];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float* cr = p.cent + (long long)(crow0 + ks[u]) * p.XP;
#pragma unroll
                    for (int i = 0; i < RECHECK_MAX_COLS_PER_LANE; ++i) {
                        const int j = lane + 32 * i;
                        cv[u][i] = j < p.N ? cr[j] : 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int i = 0; i < RECHECK_MAX_COLS_PER_LANE; ++i) {
                        if (p.metric) { const double d = xv[i] - (double)cv[u][i]; s[u] = fma(d, d, s[u]); }
                        else s[u] = fma(xv[i], (double)cv[u][i], s[u]);
                    }
            } else {
                for (int u = 0; u < U; ++u) {
                    const float* cr = p.cent + (long long)(crow0 + ks[u]) * p.XP;
                    for (int j = lane; j < p.N; j += 32) {
                        if (p.metric) { const double d = (double)xr[j] - (double)cr[j]; s[u] = fma(d, d, s[u]); }
                        else s[u] = fma((double)xr[j], (double)cr[j], s[u]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) s[u] = warp_sum(s[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (u < n) {
                    const int k = ks[u];
                    const double d = metric_distance(p.metric, s[u], xn, p.cnorm[crow0 + k]);
                    if (d != d && !any_nan) { any_nan = true; bd = -INFINITY; bk = k; }   // numpy argmin: first NaN wins
                    if (d < bd) { bd = d; bk = k; }
                }
            }
        };
        while (mask) {
            if (__popcll(mask) <= 2) batch(std::integral_constant<int, 2>());
            else batch(std::integral_constant<int, 4>());
        }
        if (p.labels_out) {
            if (lane == 0) {
                p.labels_out[(long long)run * p.M + row] = bk;
                atomic_add_ll(p.state + (long long)run * kStateWords + LEIDA_KM_ST_RECHECKED, 1);
                // watch on the error bound: the fast argmax lost although its margin was ratio * threshold
                if (fast_k >= 0 && fast_k != bk)
                    atomicMax(p.state + (long long)run * kStateWords + LEIDA_KM_ST_TAU_WATCH, (long long)ratio_bits);
            }
            return;
        }
        int32_t* lab = p.labels + (long long)run * p.M;
        const int old_l = lab[row];
        __syncwarp();
        if (old_l != bk) {
            if (lane == 0) {
                lab[row] = bk;
                atomic_add_ll(p.state + (long long)run * kStateWords + LEIDA_KM_ST_CHANGED, 1);
            }
            move_row(p, row, crow0, old_l, bk, lane, 32);
        }
        if (lane == 0) atomic_add_ll(p.state + (long long)run * kStateWords + LEIDA_KM_ST_RECHECKED, 1);
    };
    const int4* entries = reinterpret_cast<const int4*>(p.worklist + 8);
    for (long long e = warp; e < count; e += nwarps) {
        const int4 a = entries[2 * e], b = entries[2 * e + 1];
        resolve(a.x, a.y, (unsigned long long)(uint32_t)b.x | ((unsigned long long)(uint32_t)b.y << 32), a.z, a.w);
    }
    // Queue overflow (tensor-core flow only; its fast pass stores -2 for every undecided pair whether or
    // not the pair still fitted in the queue): find the unlisted pairs by scanning the labels of the
    // active runs and evaluate all K clusters for them.  A pair another warp is resolving from the queue right
    // now may be resolved twice -- same inputs, same label.
    if (p.labels_out && queued > p.worklist_cap) {
        const long long blocks = (p.M + 31) / 32;
        for (long long b = warp; b < (long long)p.n_runs * blocks; b += nwarps) {
            const int run = (int)(b / blocks);
            if (p.state[(long long)run * kStateWords + LEIDA_KM_ST_ACTIVE] == 0) continue;
            const long long row0 = (b - run * blocks) * 32;
            const bool mine = row0 + lane < p.M && p.labels_out[(long long)run * p.M + row0 + lane] == -2;
            unsigned todo = __ballot_sync(0xffffffffu, mine);
            while (todo) {
                const int bit = __ffs(todo) - 1;
                todo &= todo - 1;
                resolve(run, row0 + bit, 0ull, -1, 0);
            }
        }
    }
    // the last CTA to finish clears the list header for the next assign launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int ticket = atomicAdd(&p.worklist[1], 1);
        if (ticket == (int)gridDim.x - 1) {
            p.worklist[0] = 0;
            p.worklist[1] = 0;
        }
    }
}

__global__ void k_row_norms(const float* __restrict__ X, long long M, int N, long long XP, double* __restrict__ xnorm) {
    // warp per 32 rows: lanes own rows, sequential fp64 sum of squares (reference order).
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    const float* xr = X + row * XP;
    double s = 0.0;
    for (int j = 0; j < N; ++j) {
        const double v = (double)xr[j];
        s = __dadd_rn(s, __dmul_rn(v, v));
    }
    xnorm[row] = sqrt(s);
}

__global__ void k_pack_rows(const float* __restrict__ src, long long M, int N, long long sp, float* __restrict__ X,
                            long long XP) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * XP) return;
    const long long r = i / XP;
    const int j = (int)(i - r * XP);
    X[i] = j < N ? src[r * sp + j] : 0.f;
}

// One warp per centroid row: 2-norm in fp64 (fixed lane-strided order).
__device__ __forceinline__ double cent_norm_warp(const float* c, int N, int lane) {
    double s = 0.0;
    for (int j = lane; j < N; j += 32) {
        const double v = (double)c[j];
        s += v * v;
    }
    return sqrt(warp_sum(s));
}

__global__ void k_init(const float* __restrict__ Xi, long long xi_pitch, const long long* __restrict__ init_rows,
                       int n_runs, const int32_t* __restrict__ run_k, const int32_t* __restrict__ run_crow0, int N,
                       int XP, float* cent, double* cnorm, long long* acc, long long* state) {
    // one CTA per run
    const int run = blockIdx.x;
    const int K = run_k[run], crow0 = run_crow0[run];
    const int AP = XP + 8;
    for (int i = threadIdx.x; i < K * XP; i += blockDim.x) {
        const int k = i / XP, j = i - k * XP;
        cent[(long long)(crow0 + k) * XP + j] = j < N ? Xi[init_rows[crow0 + k] * xi_pitch + j] : 0.f;
    }
    for (int i = threadIdx.x; i < K * AP; i += blockDim.x) acc[(long long)crow0 * AP + i] = 0;
    if (threadIdx.x < kStateWords) {
        long long v = 0;
        if (threadIdx.x == LEIDA_KM_ST_ACTIVE) v = 1;
        if (threadIdx.x == LEIDA_KM_ST_K) v = K;
        state[(long long)run * kStateWords + threadIdx.x] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = w; k < K; k += nw) {
        const double n = cent_norm_warp(cent + (long long)(crow0 + k) * XP, N, lane);
        if (lane == 0) cnorm[crow0 + k] = n;
    }
}

__global__ void k_fill_i32(int32_t* p, long long n, int32_t v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = v;
}

// Control + centroid update, one CTA per run.
// Where k_update reads the job-wide sums from:
//   * one buffer (single GPU, or after an NCCL all-reduce);
//   * the exchange buffers of all ranks in symmetric memory, summed here (int64: exact, any order) -- the all-reduce
//     fused into its consumer, and only for the runs that are still active: either ONE multimem.ld_reduce per word
//     through the NVSwitch multicast mapping (the switch adds the ranks' words), or one load per rank over NVLink.
// With an exchange the kernel first waits until every rank has published this pass (flags written by
// k_exchange_publish on the peers, see there).
struct ReducedSource {
    const long long* acc;             // n_peers == 0
    const long long* state;
    const long long* const* peers;    // n_peers > 0: device array of every rank's buffer (acc words, then state words)
    int n_peers;
    long long state_off;
    const long long* mc;              // multicast address of the same buffer (multimem.ld_reduce), or null
    const unsigned long long* flags;  // this rank's flag words, one per rank, or null (the caller ordered the ranks)
    const unsigned long long* seq;    // device counter: passes consumed so far; this pass waits for flags >= *seq + 1
    __device__ __forceinline__ long long sum(long long word) const {
        if (mc) {
            unsigned long long v;
            asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(v) : "l"(mc + word) : "memory");
            return (long long)v;
        }
        long long v = 0;
        for (int p = 0; p < n_peers; ++p) v += *reinterpret_cast<const volatile long long*>(peers[p] + word);
        return v;
    }
    __device__ __forceinline__ long long acc_word(long long i) const { return n_peers ? sum(i) : acc[i]; }
    __device__ __forceinline__ long long state_word(long long i) const { return n_peers ? sum(state_off + i) : state[i]; }
    // four words at once: every load is issued before the first one is used (NVLink round trips overlap)
    __device__ __forceinline__ void acc_words4(const long long* idx, long long* out) const {
        if (!n_peers) {
#pragma unroll
            for (int u = 0; u < 4; ++u) out[u] = acc[idx[u]];
            return;
        }
        if (mc) {
            unsigned long long v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(v[u]) : "l"(mc + idx[u]) : "memory");
#pragma unroll
            for (int u = 0; u < 4; ++u) out[u] = (long long)v[u];
            return;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) out[u] = 0;
        for (int p0 = 0; p0 < n_peers; p0 += 4) {
            long long v[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const long long* base = peers[min(p0 + q, n_peers - 1)];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[q][u] = *reinterpret_cast<const volatile long long*>(base + idx[u]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (p0 + q < n_peers)
#pragma unroll
                    for (int u = 0; u < 4; ++u) out[u] += v[q][u];
        }
    }
    // block until every rank's words of this pass are visible (called by all threads of the CTA)
    __device__ __forceinline__ void wait_ranks() const {
        if (flags == nullptr) return;
        if ((int)threadIdx.x < n_peers) {
            const unsigned long long want = *seq + 1ull;
            const long long t0 = clock64();
            unsigned long long f;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(flags + threadIdx.x) : "memory");
                if (f < want && clock64() - t0 > (20ll << 30)) __trap();     // ~10 s: a rank died; fail loudly, do not hang
            } while (f < want);
        }
        __syncthreads();
    }
};

// Per-pass publish of this rank's partial sums for the exchange above: the acc rows of the runs that are still
// active and every run's state words are copied into this pass's half of the rank's symmetric-memory buffer (same
// word offsets as the source, so readers index it like the local buffer); then the last CTA tells every rank (this
// one included) that the rank's words of pass *seq + 1 are complete by writing that number into the rank's slot of
// each rank's flag array (release at system scope, after the copies were fenced at system scope).  Two halves
// alternate: a rank that is one pass ahead writes the other half, and it cannot be two passes ahead because its next
// k_update waits for this rank's next publish.
__global__ void __launch_bounds__(256)
k_exchange_publish(int n_runs, const int32_t* __restrict__ run_k, const int32_t* __restrict__ run_crow0, int AP,
                   const long long* __restrict__ src, long long n_acc_words, long long* __restrict__ dst,
                   unsigned long long* const* __restrict__ peer_flags, int n_peers, int my_rank,
                   unsigned long long* seq, int* counter) {
    const int run = blockIdx.x;
    const long long* st = src + n_acc_words + (long long)run * kStateWords;
    if (threadIdx.x < kStateWords) dst[n_acc_words + (long long)run * kStateWords + threadIdx.x] = st[threadIdx.x];
    if (st[LEIDA_KM_ST_ACTIVE] != 0) {
        const long long w0 = (long long)run_crow0[run] * AP, n = (long long)run_k[run] * AP;
        const longlong2* s2 = reinterpret_cast<const longlong2*>(src + w0);      // AP is a multiple of 8 words
        longlong2* d2 = reinterpret_cast<longlong2*>(dst + w0);
        for (long long i = threadIdx.x; i < n / 2; i += blockDim.x) d2[i] = s2[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int ticket = atomicAdd(counter, 1);
        if (ticket == n_runs - 1) {
            *counter = 0;
            __threadfence_system();
            const unsigned long long v = *seq + 1ull;
            *seq = v;
            for (int p = 0; p < n_peers; ++p)
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flags[p] + my_rank), "l"(v) : "memory");
        }
    }
}

__global__ void __launch_bounds__(256)
k_update(int n_runs, const int32_t* __restrict__ run_k, const int32_t* __restrict__ run_crow0, int N, int XP,
         const ReducedSource red, float* cent, double* cnorm,
         long long* state, int n_iter, int32_t* n_active, const int32_t* __restrict__ run_pcol,
         __nv_bfloat16* __restrict__ C1, __nv_bfloat16* __restrict__ C2, int NPK, volatile int32_t* mailbox) {
    const int run = blockIdx.x;
    long long* st = state + (long long)run * kStateWords;
    __shared__ int s_go;
    __shared__ int s_empty;
    __shared__ long long s_cnt[LEIDA_KM_MAX_K];
    const int K = run_k[run], crow0 = run_crow0[run];
    const int AP = XP + 8;
    red.wait_ranks();
    const bool was_active = st[LEIDA_KM_ST_ACTIVE] != 0;        // block-uniform; only thread 0 writes it, later
    // every thread fetches one cluster's member count (and checks it for emptiness)
    int my_empty = 0;
    if (was_active && threadIdx.x < K) {
        const long long c = red.acc_word((long long)(crow0 + threadIdx.x) * AP + XP);
        s_cnt[threadIdx.x] = c;
        if (c <= 0) my_empty = 1;
    }
    const int any_empty = __syncthreads_or(my_empty);
    if (threadIdx.x == 0) {
        int go = 0;
        s_empty = 0;
        if (st[LEIDA_KM_ST_ACTIVE] != 0) {
            const long long p = st[LEIDA_KM_ST_PASSES];  // index of the pass that just finished
            st[LEIDA_KM_ST_PASSES] = p + 1;
            const long long changed = red.state_word((long long)run * kStateWords + LEIDA_KM_ST_CHANGED);
            st[LEIDA_KM_ST_CHANGED_TOTAL] += changed;
            st[LEIDA_KM_ST_RECHECK_TOTAL] += red.state_word((long long)run * kStateWords + LEIDA_KM_ST_RECHECKED);
            // pass p >= 1 is Lloyd iteration p-1; the reference tests convergence from iteration 1 on
            // (clustering/_clustering.py:129-132) and runs at most n_iter iterations (:118).
            if ((p >= 2 && changed == 0) || p >= n_iter) st[LEIDA_KM_ST_ACTIVE] = 0;
            else go = 1;
            if (go && any_empty) {
                for (int k = 0; k < K; ++k)
                    if (s_cnt[k] <= 0) { s_empty = k + 1; break; }
                st[LEIDA_KM_ST_EMPTY] = s_empty;
                st[LEIDA_KM_ST_ACTIVE] = 0;
                go = 0;
            }
            st[LEIDA_KM_ST_CHANGED] = 0;
            st[LEIDA_KM_ST_RECHECKED] = 0;
        }
        s_go = go;
        if (go) atomicAdd(n_active, 1);
    }
    __syncthreads();
    const int pv = run_pcol ? run_pcol[run] : -1;      // packed column | slot width << 24 (k_plan)
    if (s_go) {
        // four elements per thread and step: their (possibly remote) loads are all in flight together
        for (int i0 = threadIdx.x; i0 < K * XP; i0 += 4 * blockDim.x) {
            long long idx[4], w[4];
            int kk[4], jj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = min(i0 + u * (int)blockDim.x, K * XP - 1);
                kk[u] = i / XP;
                jj[u] = i - kk[u] * XP;
                idx[u] = (long long)(crow0 + kk[u]) * AP + min(jj[u], N - 1);
            }
            red.acc_words4(idx, w);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u * (int)blockDim.x < K * XP) {
                    float v = 0.f;
                    if (jj[u] < N) v = (float)(((double)w[u] * (1.0 / kFracScale)) / (double)s_cnt[kk[u]]);
                    cent[(long long)(crow0 + kk[u]) * XP + jj[u]] = v;
                }
            }
        }
        __syncthreads();
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int k = w; k < K; k += nw) {
            const double n = cent_norm_warp(cent + (long long)(crow0 + k) * XP, N, lane);
            if (lane == 0) cnorm[crow0 + k] = n;
        }
        __syncthreads();
    }
    // n_active[0] counts the runs that continue, n_active[1] the CTAs that have decided: the last
    // one publishes {active runs, passes completed} to the host-visible mailbox and re-arms both.
    if (threadIdx.x == 0) {
        __threadfence();
        const int ticket = atomicAdd(&n_active[1], 1);
        if (ticket == n_runs - 1) {
            const int na = atomicAdd(&n_active[0], 0);
            n_active[2] = na;                       // device copy of the last count
            n_active[0] = 0;
            n_active[1] = 0;
            if (red.flags) *const_cast<unsigned long long*>(red.seq) += 1ull;   // every CTA has read it (all took a ticket)
            if (mailbox) {
                // device-side copy of the counters (n_active[3] = passes done, n_active[4] = first all-stopped
                // pass), published to the host as ONE aligned 64-bit store -- a single PCIe write, so the three
                // fields are always consistent and no system fence sits on the kernel's tail:
                //   bits 0-15 active runs | bits 16-39 passes done | bits 40-63 first all-stopped pass (0 = not yet)
                const int done = n_active[3] + 1;
                n_active[3] = done;
                if (na == 0 && n_active[4] == 0) n_active[4] = done;
                const unsigned long long word = (unsigned long long)(na & 0xFFFF) | ((unsigned long long)(done & 0xFFFFFF) << 16) |
                                                ((unsigned long long)(n_active[4] & 0xFFFFFF) << 40);
                *reinterpret_cast<volatile unsigned long long*>(mailbox) = word;
            }
        }
    }
    // tensor-core operands of this run at its packed position: unit-normalised centroids, bf16 hi/lo.
    // Written for every packed run (also one that just stopped), so the packed matrix never holds
    // stale rows of another run after a re-plan.
    if (pv >= 0) {
        const int W = pv >> 24, pc = pv & 0xFFFFFF;    // the run's slot: K centroids + padding columns that score -1e30
        for (int i = threadIdx.x; i < W * NPK; i += blockDim.x) {
            const int k = i / NPK, j = i - k * NPK;
            float v = 0.f;
            if (k >= K) {
                v = j == N ? -1e30f : 0.f;
            } else if (j < N) {
                const double cn = cnorm[crow0 + k];
                const float inv = cn > 0.0 ? (float)(1.0 / cn) : 0.f;
                v = cent[(long long)(crow0 + k) * XP + j] * inv;
            }
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            C1[(long long)(pc + k) * NPK + j] = hi;
            C2[(long long)(pc + k) * NPK + j] = __float2bfloat16_rn(v - __bfloat162float(hi));
        }
    }
}

// Control + centroid update, second form: the exchange with the other GPUs and the whole update of a run happen in ONE
// kernel and the centroid rows never leave registers between the sums and the packed tensor-core operands.
//   * exchange (several GPUs): the CTA of run r copies the run's acc rows + state words into this pass's half of the
//     rank's symmetric buffer, raises flag[my_rank][r] = pass number on every rank (one system fence, release stores),
//     and waits for flag[*][r] of all ranks.  Per-run flags: no grid-wide ticket, no separate publish launch, and a run
//     is held up only by its own counterparts.  Every word is then read as the sum over ranks (multimem.ld_reduce
//     through the multicast mapping, else one load per rank).
//   * update: a warp owns whole centroid rows: its lanes hold the row's sums (all loads of a row in flight together),
//     form the means, reduce the norm with shuffles (same summation order as cent_norm_warp) and write the float32
//     centroid, the norm and the unit-normalised bf16 hi/lo operands straight from registers: one round trip to memory
//     instead of four dependent ones.
// Padding rows / columns of a slot are constant after leida_kmeans_tc_plan(pack = 1) and are not rewritten; a run
// that stops keeps its last operands until the next re-plan (its labels are ignored from then on).
struct ExchangeParams {
    long long* my_half;                       // null: single GPU (sums are read from acc / state directly)
    const long long* const* peers;            // every rank's half of this pass
    const long long* mc;                      // multicast address of the half, or null
    unsigned long long* const* peer_flags;    // every rank's flag array, [n_ranks][flag_pitch]
    const unsigned long long* flags_local;
    unsigned long long* seq;                  // passes consumed so far (device counter)
    int n_peers, my_rank, flag_pitch;
    long long state_off;
};

template <int CPL>
__global__ void __launch_bounds__(256)
k_update_fused(int n_runs, const int32_t* __restrict__ run_k, const int32_t* __restrict__ run_crow0, int N, int XP,
               const ExchangeParams x, long long* local, float* cent, double* cnorm, long long* state, int n_iter,
               int32_t* n_active, const int32_t* __restrict__ run_pcol, __nv_bfloat16* __restrict__ C1,
               __nv_bfloat16* __restrict__ C2, int NPK, volatile int32_t* mailbox) {
    const int run = blockIdx.x;
    long long* st = state + (long long)run * kStateWords;
    __shared__ int s_go;
    __shared__ long long s_cnt[LEIDA_KM_MAX_K];
    const int K = run_k[run], crow0 = run_crow0[run];
    const int AP = XP + 8;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool was_active = st[LEIDA_KM_ST_ACTIVE] != 0;        // block-uniform; only thread 0 writes it, later
    const bool xchg = x.my_half != nullptr;
    const long long acc0 = (long long)crow0 * AP;
    if (xchg && was_active) {
        const unsigned long long want = *x.seq + 1ull;
        const longlong2* s2 = reinterpret_cast<const longlong2*>(local + acc0);       // AP is a multiple of 8 words
        longlong2* d2 = reinterpret_cast<longlong2*>(x.my_half + acc0);
        for (int i = threadIdx.x; i < K * AP / 2; i += blockDim.x) d2[i] = s2[i];
        if (threadIdx.x < kStateWords)
            x.my_half[x.state_off + (long long)run * kStateWords + threadIdx.x] =
                local[x.state_off + (long long)run * kStateWords + threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0) {
            // release pattern: ONE system-scope fence (the CTA's copies are ordered before it by the barrier, and fences
            // are cumulative), then a relaxed flag store per rank -- not a release store per rank, which is a fence each
            __threadfence_system();
            for (int p = 0; p < x.n_peers; ++p)
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;"
                             ::"l"(x.peer_flags[p] + (long long)x.my_rank * x.flag_pitch + run), "l"(want) : "memory");
        }
        if ((int)threadIdx.x < x.n_peers) {
            const unsigned long long* f = x.flags_local + (long long)threadIdx.x * x.flag_pitch + run;
            const long long t0 = clock64();
            unsigned long long v;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
                if (v < want && clock64() - t0 > (20ll << 30)) __trap();     // ~10 s: a rank died; fail loudly
            } while (v < want);
        }
        __syncthreads();
    }
    auto sum_word = [&](long long idx) -> long long {
        if (!xchg) return local[idx];
        if (x.mc) {
            unsigned long long v;
            asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(v) : "l"(x.mc + idx) : "memory");
            return (long long)v;
        }
        long long v = 0;
        for (int p = 0; p < x.n_peers; ++p) v += *reinterpret_cast<const volatile long long*>(x.peers[p] + idx);
        return v;
    };
    // The sums of a row: all CPL loads of a lane are issued before any is used; RB rows per warp are fetched together,
    // and the first batch is requested BEFORE the counts and the continue / stop decision come back (the addresses do
    // not depend on them), so a run that continues pays one memory round trip, not three.
    constexpr int RB = CPL <= 8 ? 4 : CPL <= 13 ? 3 : CPL <= 16 ? 2 : 1;
    auto load_row = [&](int k, long long* wv) {
        const long long row = acc0 + (long long)k * AP;
        if (!xchg) {
#pragma unroll
            for (int c = 0; c < CPL; ++c) { const int j = lane + 32 * c; wv[c] = j < N ? local[row + j] : 0; }
        } else if (x.mc) {
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int j = lane + 32 * c;
                unsigned long long v = 0;
                if (j < N)
                    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(v) : "l"(x.mc + row + j) : "memory");
                wv[c] = (long long)v;
            }
        } else {
#pragma unroll
            for (int c = 0; c < CPL; ++c) wv[c] = 0;
            for (int p = 0; p < x.n_peers; ++p) {
                const long long* base = x.peers[p] + row;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int j = lane + 32 * c;
                    if (j < N) wv[c] += *reinterpret_cast<const volatile long long*>(base + j);
                }
            }
        }
    };
    long long wv[RB][CPL];
    if (was_active) {
#pragma unroll
        for (int r = 0; r < RB; ++r)
            if (w + 8 * r < K) load_row(w + 8 * r, wv[r]);
    }
    int my_empty = 0;
    if (was_active && threadIdx.x < K) {
        const long long c = sum_word(acc0 + (long long)threadIdx.x * AP + XP);
        s_cnt[threadIdx.x] = c;
        if (c <= 0) my_empty = 1;
    }
    long long changed = 0, rechecked = 0;
    if (was_active && threadIdx.x == 0) {
        changed = sum_word(x.state_off + (long long)run * kStateWords + LEIDA_KM_ST_CHANGED);
        rechecked = sum_word(x.state_off + (long long)run * kStateWords + LEIDA_KM_ST_RECHECKED);
    }
    const int any_empty = __syncthreads_or(my_empty);
    if (threadIdx.x == 0) {
        int go = 0;
        if (was_active) {
            const long long p = st[LEIDA_KM_ST_PASSES];  // index of the pass that just finished
            st[LEIDA_KM_ST_PASSES] = p + 1;
            st[LEIDA_KM_ST_CHANGED_TOTAL] += changed;
            st[LEIDA_KM_ST_RECHECK_TOTAL] += rechecked;
            // pass p >= 1 is Lloyd iteration p-1; the reference tests convergence from iteration 1 on
            // (clustering/_clustering.py:129-132) and runs at most n_iter iterations (:118).
            if ((p >= 2 && changed == 0) || p >= n_iter) st[LEIDA_KM_ST_ACTIVE] = 0;
            else go = 1;
            if (go && any_empty) {
                int e = 0;
                for (int k = 0; k < K; ++k)
                    if (s_cnt[k] <= 0) { e = k + 1; break; }
                st[LEIDA_KM_ST_EMPTY] = e;
                st[LEIDA_KM_ST_ACTIVE] = 0;
                go = 0;
            }
            st[LEIDA_KM_ST_CHANGED] = 0;
            st[LEIDA_KM_ST_RECHECKED] = 0;
        }
        s_go = go;
        if (go) atomicAdd(n_active, 1);
    }
    __syncthreads();
    if (s_go) {
        const int pv = run_pcol ? run_pcol[run] : -1;      // packed column | slot width << 24 (k_plan)
        const int pc = pv & 0xFFFFFF;
        for (int k0 = w; k0 < K; k0 += 8 * RB) {
            if (k0 != w) {
#pragma unroll
                for (int r = 0; r < RB; ++r)
                    if (k0 + 8 * r < K) load_row(k0 + 8 * r, wv[r]);
            }
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int k = k0 + 8 * r;
                if (k >= K) break;
                const double cnt = (double)s_cnt[k];
                float v[CPL];
                double ss = 0.0;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int j = lane + 32 * c;
                    v[c] = j < N ? (float)(((double)wv[r][c] * (1.0 / kFracScale)) / cnt) : 0.f;
                    if (j < N) { const double d = (double)v[c]; ss += d * d; }
                    if (j < XP) cent[(long long)(crow0 + k) * XP + j] = v[c];
                }
                const double nrm = sqrt(warp_sum(ss));
                if (lane == 0) cnorm[crow0 + k] = nrm;
                if (pv >= 0) {
                    const float inv = nrm > 0.0 ? (float)(1.0 / nrm) : 0.f;
#pragma unroll
                    for (int c = 0; c < CPL; ++c) {
                        const int j = lane + 32 * c;
                        if (j < XP && j < NPK) {
                            const float u = v[c] * inv;                      // 0 for the bias column and the padding
                            const __nv_bfloat16 hi = __float2bfloat16_rn(u);
                            C1[(long long)(pc + k) * NPK + j] = hi;
                            C2[(long long)(pc + k) * NPK + j] = __float2bfloat16_rn(u - __bfloat162float(hi));
                        }
                    }
                }
            }
        }
    }
    // n_active[0] counts the runs that continue, n_active[1] the CTAs that have decided: the last one publishes
    // {active runs, passes completed} to the host-visible mailbox and re-arms both (see k_update).
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int ticket = atomicAdd(&n_active[1], 1);
        if (ticket == n_runs - 1) {
            const int na = atomicAdd(&n_active[0], 0);
            n_active[2] = na;
            n_active[0] = 0;
            n_active[1] = 0;
            if (xchg) *x.seq += 1ull;               // every CTA has read it (all took a ticket)
            if (mailbox) {
                const int done = n_active[3] + 1;
                n_active[3] = done;
                if (na == 0 && n_active[4] == 0) n_active[4] = done;
                const unsigned long long word = (unsigned long long)(na & 0xFFFF) | ((unsigned long long)(done & 0xFFFFFF) << 16) |
                                                ((unsigned long long)(n_active[4] & 0xFFFFFF) << 40);
                *reinterpret_cast<volatile unsigned long long*>(mailbox) = word;
            }
        }
    }
}

// Distortion of every run: sum over rows of (1 - cos(row, assigned centroid))^2, fp64 per row,
// accumulated as exact integers (2^-40 and 2^-80 words) so the result does not depend on the order
// or on the number of GPUs.  One CTA = 32 rows x all runs: the rows are read ONCE (each warp keeps its
// row in registers as doubles, lanes stride the columns), the labels of the 32 rows for a chunk of
// runs are staged in shared memory with coalesced loads, and four runs are in flight per warp.
constexpr int DIST_ROWS = 32;
constexpr int DIST_RUN_CHUNK = 256;
template <int CPL>
__global__ void __launch_bounds__(256)
k_distortion(const float* __restrict__ X, const double* __restrict__ xnorm, long long M, int N, int XP, int n_runs,
             const int32_t* __restrict__ run_crow0, const float* __restrict__ cent,
             const double* __restrict__ cnorm, const int32_t* __restrict__ labels, long long* state, int metric) {
    __shared__ int s_lab[DIST_RUN_CHUNK][DIST_ROWS];
    __shared__ unsigned long long s_hi[DIST_RUN_CHUNK], s_lo[DIST_RUN_CHUNK];
    __shared__ int s_c0[DIST_RUN_CHUNK];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long row0 = (long long)blockIdx.x * DIST_ROWS;
    constexpr int DIST_MAX_COLS_PER_LANE = CPL > 0 ? CPL : 1;
    constexpr bool in_regs = CPL > 0;
    for (int r0 = 0; r0 < n_runs; r0 += DIST_RUN_CHUNK) {
        const int nr = min(DIST_RUN_CHUNK, n_runs - r0);
        __syncthreads();
        for (int i = threadIdx.x; i < nr * DIST_ROWS; i += 256) {
            const int r = i >> 5, c = i & 31;
            const long long row = row0 + c;
            s_lab[r][c] = row < M ? labels[(long long)(r0 + r) * M + row] : -1;
        }
        for (int i = threadIdx.x; i < nr; i += 256) { s_hi[i] = 0; s_lo[i] = 0; s_c0[i] = run_crow0[r0 + i]; }
        __syncthreads();
        for (int rr = w; rr < DIST_ROWS; rr += 8) {
            const long long row = row0 + rr;
            if (row >= M) break;
            const float* xr = X + row * XP;
            const double xn = xnorm[row];
            double xv[DIST_MAX_COLS_PER_LANE];
            if (in_regs) {
#pragma unroll
                for (int i = 0; i < DIST_MAX_COLS_PER_LANE; ++i) {
                    const int j = lane + 32 * i;
                    xv[i] = j < N ? (double)xr[j] : 0.0;
                }
            }
            for (int g0 = 0; g0 < nr; g0 += 32) {
                // 32 runs per group: lane u ends up holding the dot product of run g0+u, so the fp64
                // division, the squaring and the shared-memory accumulation use all 32 lanes.
                double mine = 0.0;
                int mycrow = 0;
                const int gend = min(g0 + 32, nr);
                for (int q0 = g0; q0 < gend; q0 += 4) {
                    double s4[4] = {0.0, 0.0, 0.0, 0.0};
                    int crow[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = min(q0 + u, nr - 1);
                        const int l = s_lab[r][rr];
                        crow[u] = s_c0[r] + (l < 0 ? 0 : l);
                    }
                    if (in_regs) {
                        float cv[4][DIST_MAX_COLS_PER_LANE];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float* cr = cent + (long long)crow[u] * XP;
#pragma unroll
                            for (int i = 0; i < DIST_MAX_COLS_PER_LANE; ++i) {
                                const int j = lane + 32 * i;
                                cv[u][i] = j < N ? cr[j] : 0.f;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int i = 0; i < DIST_MAX_COLS_PER_LANE; ++i) {
                                if (metric) { const double d = xv[i] - (double)cv[u][i]; s4[u] = fma(d, d, s4[u]); }
                                else s4[u] = fma(xv[i], (double)cv[u][i], s4[u]);
                            }
                    } else {
                        for (int u = 0; u < 4; ++u) {
                            const float* cr = cent + (long long)crow[u] * XP;
                            for (int j = lane; j < N; j += 32) {
                                if (metric) { const double d = (double)xr[j] - (double)cr[j]; s4[u] = fma(d, d, s4[u]); }
                                else s4[u] = fma((double)xr[j], (double)cr[j], s4[u]);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const double t = warp_sum(s4[u]);
                        if (lane == q0 - g0 + u) { mine = t; mycrow = crow[u]; }
                    }
                }
                if (g0 + lane < gend) {
                    const double d = metric_distance(metric, mine, xn, cnorm[mycrow]);
                    const double d2 = d * d * kFracScale;   // exact scaling
                    if (d2 == d2) {
                        const double f = floor(d2);
                        atomicAdd(&s_hi[g0 + lane], (unsigned long long)(long long)f);
                        atomicAdd(&s_lo[g0 + lane], (unsigned long long)(long long)floor((d2 - f) * kFracScale));
                    } else {
                        atomicAdd(&s_hi[g0 + lane], 0x4000000000000000ULL);   // poison: NaN distance
                    }
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nr; i += 256) {
            if (s_hi[i] | s_lo[i]) {
                atomic_add_ll(state + (long long)(r0 + i) * kStateWords + LEIDA_KM_ST_DIST_HI, (long long)s_hi[i]);
                atomic_add_ll(state + (long long)(r0 + i) * kStateWords + LEIDA_KM_ST_DIST_LO, (long long)s_lo[i]);
            }
        }
    }
}

// Distortion, second form (rows up to 512 columns): the L2 traffic of the kernel above is one centroid row per (row, run)
// pair -- 546 GB at 1.2 M rows x 285 runs x 400 columns, which is what bounds it.  Here a CTA owns a tile of rows held
// in REGISTERS as doubles (12 warps x Q quads of 4 rows, lanes stride the columns) and walks over all runs; each run's
// K centroid rows are staged once per tile in shared memory (cp.async ring, NST runs deep, one barrier per run) and
// every (row, run) dot product reads its centroid from there: L2 traffic / 4-5, and when the four rows of a quad carry
// the same label (consecutive volumes of a subject usually do -- that persistence is what dwell times measure) one
// shared-memory read and one float->double conversion serve all four.  Per-row arithmetic (lane-strided fp64 FMA
// chain, fixed butterfly) does not depend on the neighbours, so results are independent of tiling and sharding.
// Persistent CTAs keep the per-run sums in shared memory and issue their global atomics once.
constexpr int D2_WARPS = 12;
constexpr int D2_THREADS = 32 * D2_WARPS;

template <int CPL, int Q, int NST>
__global__ void __launch_bounds__(D2_THREADS, 1)
k_distortion_tiled(const float* __restrict__ X, const double* __restrict__ xnorm, long long M, int XP, int n_runs,
                   const int32_t* __restrict__ run_k, const int32_t* __restrict__ run_crow0,
                   const float* __restrict__ cent, const double* __restrict__ cnorm, const int32_t* __restrict__ labels,
                   long long* state, int kmax) {
    constexpr int RPW = 4 * Q;                       // rows per warp
    constexpr int ROWS = D2_WARPS * RPW;             // rows per tile
    extern __shared__ __align__(16) unsigned char d2raw[];
    float* cbuf = reinterpret_cast<float*>(d2raw);                                   // [NST][kmax * XP]
    double* cnb = reinterpret_cast<double*>(cbuf + (size_t)NST * kmax * XP);         // [NST][kmax]
    unsigned long long* s_hi = reinterpret_cast<unsigned long long*>(cnb + NST * kmax);   // [n_runs]
    unsigned long long* s_lo = s_hi + n_runs;
    int* s_k = reinterpret_cast<int*>(s_lo + n_runs);                                // [n_runs]
    int* s_c0 = s_k + n_runs;                                                        // [n_runs]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < n_runs; i += D2_THREADS) {
        s_hi[i] = 0; s_lo[i] = 0; s_k[i] = run_k[i]; s_c0[i] = run_crow0[i];
    }
    __syncthreads();
    const long long n_tiles = (M + ROWS - 1) / ROWS;

    auto stage = [&](int run) {                      // one cp.async group per call, also when there is nothing to copy
        if (run < n_runs) {
            const int b = run % NST;
            const int K = s_k[run], crow0 = s_c0[run];
            const float* src = cent + (long long)crow0 * XP;
            float* dst = cbuf + (size_t)b * kmax * XP;
            const int chunks = K * (XP / 4);
            for (int i = threadIdx.x; i < chunks; i += D2_THREADS) cp_async16(dst + 4 * i, src + 4 * (long long)i, true);
            if (threadIdx.x < K) cnb[b * kmax + threadIdx.x] = cnorm[crow0 + threadIdx.x];
        }
        cp_async_commit();
    };

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * ROWS + (long long)w * RPW;
        double xv[RPW][CPL];
        double xn[RPW];
#pragma unroll
        for (int u = 0; u < RPW; ++u) {
            const long long row = row0 + u;
            const bool ok = row < M;
            const float* xr = X + (ok ? row : 0) * XP;
            xn[u] = ok ? xnorm[row] : 1.0;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int j = lane + 32 * i;
                xv[u][i] = (ok && j < XP) ? (double)xr[j] : 0.0;
            }
        }
        // labels of the warp's rows: lane u < RPW holds row u's label of the current run; two runs are prefetched
        const bool lab_lane = lane < RPW && row0 + lane < M;
        const int32_t* lab_ptr = labels + row0 + lane;
        int lab_cur = lab_lane ? lab_ptr[0] : 0;
        int lab_n1 = (lab_lane && 1 < n_runs) ? lab_ptr[(long long)M] : 0;
        int lab_n2 = (lab_lane && 2 < n_runs) ? lab_ptr[2 * (long long)M] : 0;
        const bool b4 = lane & 16, b3 = lane & 8;
        const int u_mine = (b4 ? 2 : 0) + (b3 ? 1 : 0);          // the row of a quad this lane finishes
        double xn_q[Q], pend[Q], pend_cn[Q];
#pragma unroll
        for (int qd = 0; qd < Q; ++qd) {
            xn_q[qd] = xn[4 * qd];
#pragma unroll
            for (int u = 1; u < 4; ++u) if (u_mine == u) xn_q[qd] = xn[4 * qd + u];
            pend[qd] = 0.0;
            pend_cn[qd] = 1.0;
        }
        __syncthreads();                       // the previous tile's last compute is done: the ring may be refilled
#pragma unroll
        for (int r = 0; r < NST - 1; ++r) stage(r);
        for (int r = 0; r < n_runs; ++r) {
            const int b = r % NST;
            cp_async_wait<NST - 2>();          // this thread's copies of run r have landed
            __syncthreads();                   // ... everyone's have, and everyone finished computing run r - 1
            stage(r + NST - 1);                // refills the buffer run r - 1 used
            const int lab_n3 = (lab_lane && r + 3 < n_runs) ? lab_ptr[(long long)(r + 3) * M] : 0;
            const float* cb = cbuf + (size_t)b * kmax * XP;
#pragma unroll
            for (int qd = 0; qd < Q; ++qd) {
                int l[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    l[u] = __shfl_sync(0xffffffffu, lab_cur, 4 * qd + u);
                    if (l[u] < 0) l[u] = 0;
                }
                double s[4] = {0.0, 0.0, 0.0, 0.0};
                if (l[0] == l[1] && l[1] == l[2] && l[2] == l[3]) {
                    const float* c0 = cb + (size_t)l[0] * XP;
#pragma unroll
                    for (int i = 0; i < CPL; ++i) {
                        const int j = lane + 32 * i;
                        const double cv = j < XP ? (double)c0[j] : 0.0;
#pragma unroll
                        for (int u = 0; u < 4; ++u) s[u] = fma(xv[4 * qd + u][i], cv, s[u]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < CPL; ++i) {
                        const int j = lane + 32 * i;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const double cv = j < XP ? (double)cb[(size_t)l[u] * XP + j] : 0.0;
                            s[u] = fma(xv[4 * qd + u][i], cv, s[u]);
                        }
                    }
                }
                // four sums over 32 lanes with 6 shuffles: halve the values held per lane twice, then a plain butterfly;
                // afterwards the 8 lanes with the same (bit 4, bit 3) all hold the sum of row u_mine of the quad
                double a0 = b4 ? s[2] : s[0], a1 = b4 ? s[3] : s[1];
                a0 += __shfl_xor_sync(0xffffffffu, b4 ? s[0] : s[2], 16);
                a1 += __shfl_xor_sync(0xffffffffu, b4 ? s[1] : s[3], 16);
                double t = b3 ? a1 : a0;
                t += __shfl_xor_sync(0xffffffffu, b3 ? a0 : a1, 8);
                t += __shfl_xor_sync(0xffffffffu, t, 4);
                t += __shfl_xor_sync(0xffffffffu, t, 2);
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                // lane (row u_mine, slot r & 7) keeps the dot product and the centroid norm: the division, the squaring
                // and the exact-integer conversion are done for eight runs at a time with all 32 lanes busy
                int l_mine = l[0];
#pragma unroll
                for (int u = 1; u < 4; ++u) if (u_mine == u) l_mine = l[u];
                const double cn_mine = cnb[b * kmax + l_mine];
                if ((lane & 7) == (r & 7)) { pend[qd] = t; pend_cn[qd] = cn_mine; }
            }
            if ((r & 7) == 7 || r == n_runs - 1) {
                const int my_run = (r & ~7) + (lane & 7);
                unsigned long long hi = 0, lo = 0;
#pragma unroll
                for (int qd = 0; qd < Q; ++qd) {
                    if (my_run <= r && row0 + 4 * qd + u_mine < M) {
                        const double d = cosine_distance(pend[qd], xn_q[qd], pend_cn[qd]);
                        const double d2 = d * d * kFracScale;   // exact scaling
                        if (d2 == d2) {
                            const double f = floor(d2);
                            hi += (unsigned long long)(long long)f;
                            lo += (unsigned long long)(long long)floor((d2 - f) * kFracScale);
                        } else {
                            hi += 0x4000000000000000ULL;        // poison: NaN distance
                        }
                    }
                }
                hi += __shfl_xor_sync(0xffffffffu, hi, 8);
                lo += __shfl_xor_sync(0xffffffffu, lo, 8);
                hi += __shfl_xor_sync(0xffffffffu, hi, 16);
                lo += __shfl_xor_sync(0xffffffffu, lo, 16);
                if (lane < 8 && my_run <= r && (hi | lo)) {
                    atomicAdd(&s_hi[my_run], hi);
                    atomicAdd(&s_lo[my_run], lo);
                }
            }
            lab_cur = lab_n1; lab_n1 = lab_n2; lab_n2 = lab_n3;
        }
        cp_async_wait<0>();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_runs; i += D2_THREADS) {
        if (s_hi[i] | s_lo[i]) {
            atomic_add_ll(state + (long long)i * kStateWords + LEIDA_KM_ST_DIST_HI, (long long)s_hi[i]);
            atomic_add_ll(state + (long long)i * kStateWords + LEIDA_KM_ST_DIST_LO, (long long)s_lo[i]);
        }
    }
}

// ---- mbarrier / bulk-copy wrappers of the pipelined distortion kernel --------------------------------------------
__device__ __forceinline__ uint32_t d3_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void d3_bar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(d3_smem(bar)), "r"(count));
}
__device__ __forceinline__ void d3_bar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(d3_smem(bar)) : "memory");
}
__device__ __forceinline__ void d3_bar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(d3_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void d3_bar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = d3_smem(bar);
    uint32_t done = 0;
    long long t0 = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile(
            "{\n.reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.b32 %0, 1, 0, P1;\n}\n"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && (spin & 0xFFFFu) == 0xFFFFu) {        // a lost arrival fails loudly (by time, ~20 s of SM clock)
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > (40ll << 30)) __trap();
        }
    }
}
// one contiguous block global -> shared through the bulk-copy engine; completion is counted in bytes on `bar`
__device__ __forceinline__ void d3_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d3_smem(dst)), "l"(src), "r"(bytes), "r"(d3_smem(bar)) : "memory");
}

// Distortion, third form: the tiled kernel above spends its issue slots on everything but the FMAs (ncu, config 3:
// 390 warp instructions per (quad, run), 104 of them the 52 DFMAs + their conversions; 16 % of the stall samples on
// the block barrier of every run, staging by per-thread cp.async another ~60 instructions per thread and run).  Here
//  * a run's centroids are ONE contiguous block of `cent` (K rows x XP floats): a single bulk copy issued by one
//    thread stages them (the warps take turns), completion is counted on an mbarrier (`full`), and the ring is handed
//    back through a second mbarrier (`empty`, one arrival per warp): no block barrier, warps drift apart by up to
//    NST - 2 runs, and the ring keeps running across the tiles of a persistent CTA (the staged data does not depend on
//    the tile);
//  * the column bound is tested once (only the last 32-column group can be ragged), not per element;
//  * a quad whose rows carry two distinct labels (in any arrangement: one transition inside four consecutive volumes,
//    or a volume flickering between two neighbouring states) reads two centroid rows instead of four.
// Arithmetic per row is exactly the tiled kernel's (same lane-strided fp64 FMA chain, same butterfly), so the sums are
// bit-identical.
// Tried and dropped: staging the centroids as doubles (converted once per launch) to take the float -> double
// conversions out of the loop -- 64 KB per run leaves a three-deep ring without slack, and the kernel was 8 % slower.
template <int CPL, int Q>
__global__ void __launch_bounds__(D2_THREADS, 1)
k_distortion_piped(const float* __restrict__ X, const double* __restrict__ xnorm, long long M, int XP, int n_runs,
                   const int32_t* __restrict__ run_k, const int32_t* __restrict__ run_crow0,
                   const float* __restrict__ cent, const double* __restrict__ cnorm, const int32_t* __restrict__ labels,
                   long long* state, int kmax, int NST) {
    constexpr int RPW = 4 * Q;                       // rows per warp
    constexpr int ROWS = D2_WARPS * RPW;             // rows per tile
    extern __shared__ __align__(128) unsigned char d3raw[];
    const size_t stage_floats = ((size_t)kmax * XP + 31) / 32 * 32;                  // 128-byte multiples
    float* cbuf = reinterpret_cast<float*>(d3raw);                                   // [NST][stage_floats] + 32 slack
    double* cnb = reinterpret_cast<double*>(cbuf + (size_t)NST * stage_floats + 32); // [NST][kmax]
    unsigned long long* s_hi = reinterpret_cast<unsigned long long*>(cnb + NST * kmax);   // [n_runs]
    unsigned long long* s_lo = s_hi + n_runs;
    uint64_t* full = reinterpret_cast<uint64_t*>(s_lo + n_runs);                     // [NST]
    uint64_t* empty = full + NST;                                                    // [NST]
    int* s_k = reinterpret_cast<int*>(empty + NST);                                  // [n_runs]
    int* s_c0 = s_k + n_runs;                                                        // [n_runs]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < n_runs; i += D2_THREADS) {
        s_hi[i] = 0; s_lo[i] = 0; s_k[i] = run_k[i]; s_c0[i] = run_crow0[i];
    }
    if (threadIdx.x == 0) {
        for (int b = 0; b < NST; ++b) { d3_bar_init(&full[b], 1); d3_bar_init(&empty[b], D2_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long n_tiles = (M + ROWS - 1) / ROWS;
    const long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total = my_tiles * n_runs;       // ring iterations of this CTA

    // the calling warp stages a run into ring buffer b
    auto stage = [&](int run, int b) {
        const int K = s_k[run], crow0 = s_c0[run];
        for (int i = lane; i < K; i += 32) cnb[b * kmax + i] = cnorm[crow0 + i];
        __syncwarp();
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)K * (uint32_t)XP * 4u;
            d3_bar_expect_tx(&full[b], bytes);
            d3_bulk_g2s(cbuf + (size_t)b * stage_floats, cent + (long long)crow0 * XP, bytes, &full[b]);
        }
    };
    // Iteration g stages iteration g + LEAD into the buffer iteration g - 1 - SLACK used: the staging warp then waits
    // for a buffer every warp left SLACK iterations ago -- with SLACK = 0 it would wait for the slowest warp of the
    // previous iteration, i.e. a block barrier in disguise (measured: a quarter of all stall samples)
    const int SLACK = NST >= 5 ? 2 : (NST == 4 ? 1 : 0);
    const int LEAD = NST - 1 - SLACK;
    if (w == 0)
        for (int g0 = 0; g0 < LEAD && g0 < total; ++g0) stage(g0 % n_runs, g0);

    // CPL == ceil(XP / 32) (the launcher picks it so): only the last column group can be ragged
    const bool last_in = lane + 32 * (CPL - 1) < XP;
    const bool b4 = lane & 16, b3 = lane & 8;
    const int u_mine = (b4 ? 2 : 0) + (b3 ? 1 : 0);          // the row of a quad this lane finishes
    // ring iteration g = (tile index of this CTA) * n_runs + r uses buffer b = g % NST in its (g / NST)-th use (parity ph)
    long long g = 0;
    int b = 0;
    uint32_t ph = 0;
    int r_ahead = LEAD % n_runs;                     // run of iteration g + LEAD
    int bt = 0;                                      // buffer (and parity pt) of iteration g - 1 - SLACK, once g > SLACK
    uint32_t pt = 0;
    int pw = 0;                                      // the warp that stages during iteration g: g % D2_WARPS -- the duty
                                                     // (a dependent global load + the wait for the slowest warp) rotates,
                                                     // a fixed producer warp would lag behind and pace the whole CTA
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * ROWS + (long long)w * RPW;
        double xv[RPW][CPL];
        double xn[RPW];
#pragma unroll
        for (int u = 0; u < RPW; ++u) {
            const long long row = row0 + u;
            const bool ok = row < M;
            const float* xr = X + (ok ? row : 0) * XP;
            xn[u] = ok ? xnorm[row] : 1.0;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int j = lane + 32 * i;
                xv[u][i] = (ok && j < XP) ? (double)xr[j] : 0.0;
            }
        }
        const bool lab_lane = lane < RPW && row0 + lane < M;
        const int32_t* lab_ptr = labels + row0 + lane;
        int lab_cur = lab_lane ? lab_ptr[0] : 0;
        int lab_n1 = (lab_lane && 1 < n_runs) ? lab_ptr[(long long)M] : 0;
        int lab_n2 = (lab_lane && 2 < n_runs) ? lab_ptr[2 * (long long)M] : 0;
        const int32_t* lab_pf = lab_ptr + 3 * (long long)M;      // labels of run r + 3
        double xn_q[Q], pend[Q], pend_cn[Q];
#pragma unroll
        for (int qd = 0; qd < Q; ++qd) {
            xn_q[qd] = xn[4 * qd];
#pragma unroll
            for (int u = 1; u < 4; ++u) if (u_mine == u) xn_q[qd] = xn[4 * qd + u];
            pend[qd] = 0.0;
            pend_cn[qd] = 1.0;
        }
        for (int r = 0; r < n_runs; ++r, ++g) {
            if (w == pw && g + LEAD < total) {
                if (g > SLACK) d3_bar_wait(&empty[bt], pt);   // every warp has left iteration g - 1 - SLACK
                stage(r_ahead, g > SLACK ? bt : (int)g + LEAD);
            }
            if (g > SLACK && ++bt == NST) { bt = 0; pt ^= 1u; }
            if (++r_ahead == n_runs) r_ahead = 0;
            if (++pw == D2_WARPS) pw = 0;
            const int lab_n3 = (lab_lane && r + 3 < n_runs) ? *lab_pf : 0;
            lab_pf += M;
            d3_bar_wait(&full[b], ph);
            const float* cb = cbuf + (size_t)b * stage_floats;
#pragma unroll
            for (int qd = 0; qd < Q; ++qd) {
                int l[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    l[u] = __shfl_sync(0xffffffffu, lab_cur, 4 * qd + u);
                    if (l[u] < 0) l[u] = 0;
                }
                double s[4] = {0.0, 0.0, 0.0, 0.0};
                // two labels in the quad (any arrangement: row 0 carries `la`, bit u - 1 of MASK says row u carries the
                // other one): two centroid rows serve the four rows; MASK = 0: one row serves all four
                auto dots = [&](auto mtag, const float* ca, const float* cq) {
                    constexpr int MASK = decltype(mtag)::value;
#pragma unroll
                    for (int i = 0; i < CPL; ++i) {
                        const int j = lane + 32 * i;
                        double va = (double)ca[j], vq = MASK ? (double)cq[j] : 0.0;
                        if (i == CPL - 1) { va = last_in ? va : 0.0; vq = last_in ? vq : 0.0; }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            s[u] = fma(xv[4 * qd + u][i], (u > 0 && ((MASK >> (u - 1)) & 1)) ? vq : va, s[u]);
                    }
                };
                const int la = l[0];
                const int lb = l[1] != la ? l[1] : (l[2] != la ? l[2] : l[3]);
                const int mask = (l[1] != la ? 1 : 0) | (l[2] != la ? 2 : 0) | (l[3] != la ? 4 : 0);
                const bool two = (l[1] == la || l[1] == lb) && (l[2] == la || l[2] == lb) && (l[3] == la || l[3] == lb);
                const float* c0 = cb + (size_t)la * XP;
                const float* c3 = cb + (size_t)l[3] * XP;
                if (two) {
                    const float* cq = cb + (size_t)lb * XP;
                    switch (mask) {
                        case 0: dots(std::integral_constant<int, 0>{}, c0, c0); break;
                        case 1: dots(std::integral_constant<int, 1>{}, c0, cq); break;
                        case 2: dots(std::integral_constant<int, 2>{}, c0, cq); break;
                        case 3: dots(std::integral_constant<int, 3>{}, c0, cq); break;
                        case 4: dots(std::integral_constant<int, 4>{}, c0, cq); break;
                        case 5: dots(std::integral_constant<int, 5>{}, c0, cq); break;
                        case 6: dots(std::integral_constant<int, 6>{}, c0, cq); break;
                        default: dots(std::integral_constant<int, 7>{}, c0, cq); break;
                    }
                }
                else {
                    // four rows, four centroid rows: two rows at a time (two independent chains with few live
                    // temporaries -- with all four in one loop the compiler runs the chains one after the other)
                    const float* cr[4] = {c0, cb + (size_t)l[1] * XP, cb + (size_t)l[2] * XP, c3};
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
#pragma unroll
                        for (int i = 0; i < CPL; ++i) {
                            const int j = lane + 32 * i;
                            double va = (double)cr[2 * h][j], vq = (double)cr[2 * h + 1][j];
                            if (i == CPL - 1) { va = last_in ? va : 0.0; vq = last_in ? vq : 0.0; }
                            s[2 * h] = fma(xv[4 * qd + 2 * h][i], va, s[2 * h]);
                            s[2 * h + 1] = fma(xv[4 * qd + 2 * h + 1][i], vq, s[2 * h + 1]);
                        }
                    }
                }
                // four sums over 32 lanes with 6 shuffles (see k_distortion_tiled)
                double a0 = b4 ? s[2] : s[0], a1 = b4 ? s[3] : s[1];
                a0 += __shfl_xor_sync(0xffffffffu, b4 ? s[0] : s[2], 16);
                a1 += __shfl_xor_sync(0xffffffffu, b4 ? s[1] : s[3], 16);
                double t = b3 ? a1 : a0;
                t += __shfl_xor_sync(0xffffffffu, b3 ? a0 : a1, 8);
                t += __shfl_xor_sync(0xffffffffu, t, 4);
                t += __shfl_xor_sync(0xffffffffu, t, 2);
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                int l_mine = l[0];
#pragma unroll
                for (int u = 1; u < 4; ++u) if (u_mine == u) l_mine = l[u];
                const double cn_mine = cnb[b * kmax + l_mine];
                if ((lane & 7) == (r & 7)) { pend[qd] = t; pend_cn[qd] = cn_mine; }
            }
            __syncwarp();                              // every lane's reads of the buffer are done
            if (lane == 0) d3_bar_arrive(&empty[b]);
            if ((r & 7) == 7 || r == n_runs - 1) {
                const int my_run = (r & ~7) + (lane & 7);
                unsigned long long hi = 0, lo = 0;
#pragma unroll
                for (int qd = 0; qd < Q; ++qd) {
                    if (my_run <= r && row0 + 4 * qd + u_mine < M) {
                        const double d = cosine_distance(pend[qd], xn_q[qd], pend_cn[qd]);
                        const double d2 = d * d * kFracScale;   // exact scaling
                        if (d2 == d2) {
                            const double f = floor(d2);
                            hi += (unsigned long long)(long long)f;
                            lo += (unsigned long long)(long long)floor((d2 - f) * kFracScale);
                        } else {
                            hi += 0x4000000000000000ULL;        // poison: NaN distance
                        }
                    }
                }
                hi += __shfl_xor_sync(0xffffffffu, hi, 8);
                lo += __shfl_xor_sync(0xffffffffu, lo, 8);
                hi += __shfl_xor_sync(0xffffffffu, hi, 16);
                lo += __shfl_xor_sync(0xffffffffu, lo, 16);
                if (lane < 8 && my_run <= r && (hi | lo)) {
                    atomicAdd(&s_hi[my_run], hi);
                    atomicAdd(&s_lo[my_run], lo);
                }
            }
            lab_cur = lab_n1; lab_n1 = lab_n2; lab_n2 = lab_n3;
            if (++b == NST) { b = 0; ph ^= 1u; }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_runs; i += D2_THREADS) {
        if (s_hi[i] | s_lo[i]) {
            atomic_add_ll(state + (long long)i * kStateWords + LEIDA_KM_ST_DIST_HI, (long long)s_hi[i]);
            atomic_add_ll(state + (long long)i * kStateWords + LEIDA_KM_ST_DIST_LO, (long long)s_lo[i]);
        }
    }
}

__global__ void k_zero_dist(long long* state, int n_runs) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_runs) {
        state[(long long)r * kStateWords + LEIDA_KM_ST_DIST_HI] = 0;
        state[(long long)r * kStateWords + LEIDA_KM_ST_DIST_LO] = 0;
    }
}

// Reference-faithful float32 means: one warp per (run, cluster, 32-column group), rows scanned in
// order, float32 adds in row order, float32 divide (numpy mean(axis=0) on float32 rows).
__global__ void __launch_bounds__(128)
k_means_f32_seq(const float* __restrict__ X, long long M, int N, int XP, const int32_t* __restrict__ run_k,
                const int32_t* __restrict__ run_crow0, const int32_t* __restrict__ labels,
                const long long* __restrict__ state, float* cent, int groups_per_row, int total_rows_dummy) {
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int run = blockIdx.y;
    if (state[(long long)run * kStateWords + LEIDA_KM_ST_ACTIVE] == 0) return;
    const int K = run_k[run], crow0 = run_crow0[run];
    if (gw >= (long long)K * groups_per_row) return;
    const int k = (int)(gw / groups_per_row);
    const int col = (int)(gw % groups_per_row) * 32 + lane;
    const int32_t* lab = labels + (long long)run * M;
    float acc = 0.f;
    long long cnt = 0;
    bool first = true;
    for (long long m0 = 0; m0 < M; m0 += 32) {
        const long long m = m0 + lane;
        const bool mine = (m < M) && (lab[m] == k);
        unsigned mask = __ballot_sync(0xffffffffu, mine);
        cnt += __popc(mask);
        while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            const float v = col < XP ? X[(m0 + b) * XP + col] : 0.f;
            if (first) { acc = v; first = false; }
            else acc = __fadd_rn(acc, v);
        }
    }
    if (col < XP) {
        float out = 0.f;
        if (col < N) out = cnt > 0 ? __fdiv_rn(acc, (float)cnt) : __int_as_float(0x7fc00000);
        cent[(long long)(crow0 + k) * XP + col] = out;
    }
}

__global__ void k_cent_norms(const int32_t* __restrict__ run_k, const int32_t* __restrict__ run_crow0,
                             const long long* __restrict__ state, const float* __restrict__ cent, int N, int XP,
                             double* cnorm) {
    const int run = blockIdx.x;
    if (state[(long long)run * kStateWords + LEIDA_KM_ST_ACTIVE] == 0) return;
    const int K = run_k[run], crow0 = run_crow0[run];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = w; k < K; k += nw) {
        const double n = cent_norm_warp(cent + (long long)(crow0 + k) * XP, N, lane);
        if (lane == 0) cnorm[crow0 + k] = n;
    }
}

// transform / predict: warp per row, fp64; metric 0 = cosine, 1 = Euclidean (scipy cdist semantics).
__global__ void __launch_bounds__(256)
k_transform(const float* __restrict__ X, long long M, int N, long long XP, const float* __restrict__ cent, int K,
            long long CP, double* dist, int32_t* labels, int metric) {
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    const float* xr = X + row * XP;
    double xx = 0.0;
    for (int j = lane; j < N; j += 32) xx += (double)xr[j] * (double)xr[j];
    const double xn = sqrt(warp_sum(xx));
    double bd = INFINITY;
    int bk = 0;
    bool any = false;
    for (int k = 0; k < K; ++k) {
        const float* cr = cent + k * CP;
        double s = 0.0, cc = 0.0;
        for (int j = lane; j < N; j += 32) {
            const double c = (double)cr[j];
            if (metric) { const double d = (double)xr[j] - c; s += d * d; }
            else s += (double)xr[j] * c;
            cc += c * c;
        }
        s = warp_sum(s);
        cc = sqrt(warp_sum(cc));
        const double d = metric_distance(metric, s, xn, cc);
        if (dist && lane == 0) dist[row * K + k] = d;
        if (d != d && !any) { /* NaN: numpy argmin returns the first NaN */ bd = -INFINITY; bk = k; any = true; }
        if (d < bd) { bd = d; bk = k; }
    }
    if (labels && lane == 0) labels[row] = bk;
}

__global__ void k_remap(int32_t* labels, long long M, const int32_t* __restrict__ lut, int K) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) {
        const int l = labels[i];
        if (l >= 0 && l < K) labels[i] = lut[l];
    }
}

// out[c][i] = lut[c][labels[src_run[c]][i]] for every selected run c: the relabelled labels of the best restart
// of every K in one launch.  grid = (ceil(M / 256), n_out).
__global__ void k_remap_batch(const int32_t* __restrict__ labels, long long M, const int32_t* __restrict__ src_run,
                              const int32_t* __restrict__ lut, int kmax, int32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (i < M) {
        const int l = labels[(long long)src_run[c] * M + i];
        out[(long long)c * M + i] = (l >= 0 && l < kmax) ? lut[c * kmax + l] : l;
    }
}

template <int KP, int RPT>
static int launch_assign(const AssignParams& p, int n_runs, cudaStream_t st) {
    constexpr int TR = 128 * RPT;
    const size_t smem = (size_t)3 * (TR * 32 + KP * 32) * sizeof(float) + KP * sizeof(float) + TR * sizeof(int);
    LEIDA_CUDA_TRY(cudaFuncSetAttribute(k_assign<KP, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((p.M + TR - 1) / TR), (unsigned)n_runs);
    k_assign<KP, RPT><<<grid, 128, smem, st>>>(p);
    return check_launch("k_assign");
}

}  // namespace leida

using namespace leida;

extern "C" int leida_kmeans_row_norms(const float* X, int64_t M, int N, int64_t x_pitch, double* xnorm,
                                      leida_stream_t stream) {
    LEIDA_REQUIRE(X && xnorm, "null pointer");
    LEIDA_REQUIRE(M >= 0 && N >= 1 && x_pitch >= N, "sizes");
    if (M == 0) return LEIDA_OK;
    k_row_norms<<<(unsigned)((M + 127) / 128), 128, 0, (cudaStream_t)stream>>>(X, M, N, x_pitch, xnorm);
    return check_launch("k_row_norms");
}

extern "C" int leida_kmeans_pack_rows(const float* src, int64_t M, int N, int64_t src_pitch, float* X, int64_t x_pitch,
                                      leida_stream_t stream) {
    LEIDA_REQUIRE(src && X, "null pointer");
    LEIDA_REQUIRE(M >= 0 && N >= 1 && src_pitch >= N && x_pitch >= N, "sizes");
    if (M == 0) return LEIDA_OK;
    const long long n = M * x_pitch;
    k_pack_rows<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, M, N, src_pitch, X, x_pitch);
    return check_launch("k_pack_rows");
}

static int check_runs(int n_runs, int N, int64_t x_pitch) {
    LEIDA_REQUIRE(n_runs >= 1 && n_runs <= 65535, "1 <= n_runs <= 65535 per call");
    LEIDA_REQUIRE(N >= 1 && x_pitch >= N && x_pitch % 32 == 0, "x_pitch multiple of 32 and >= N");
    return LEIDA_OK;
}

extern "C" int leida_kmeans_init(const float* X_init, int64_t xi_pitch, const int64_t* init_rows, int n_runs,
                                 const int32_t* run_k, const int32_t* run_crow0, int N, int64_t x_pitch, float* cent,
                                 double* cnorm, int64_t* acc, int32_t* labels, int64_t M, int64_t* state,
                                 leida_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_REQUIRE(X_init && init_rows && run_k && run_crow0 && cent && cnorm && acc && labels && state, "null pointer");
    int rc = check_runs(n_runs, N, x_pitch);
    if (rc) return rc;
    LEIDA_REQUIRE(xi_pitch >= N && M >= 0, "sizes");
    k_init<<<n_runs, 256, 0, st>>>(X_init, xi_pitch, (const long long*)init_rows, n_runs, run_k, run_crow0, N, (int)x_pitch,
                                   cent, cnorm, (long long*)acc, (long long*)state);
    rc = check_launch("k_init");
    if (rc) return rc;
    const long long n = (long long)n_runs * M;
    if (n > 0) {
        k_fill_i32<<<(unsigned)((n + 1023) / 1024 > 4096 ? 4096 : (n + 1023) / 1024), 256, 0, st>>>(labels, n, -1);
        rc = check_launch("k_fill_i32");
    }
    return rc;
}

extern "C" int leida_kmeans_assign(const float* X, const double* xnorm, int64_t M, int N, int64_t x_pitch,
                                   int n_runs, int kmax, const int32_t* run_k, const int32_t* run_crow0,
                                        const float* cent, const double* cnorm, int64_t* acc, int32_t* labels,
                                        int64_t* state, int32_t* worklist, int64_t worklist_cap, int metric,
                                        leida_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_REQUIRE(X && xnorm && run_k && run_crow0 && cent && cnorm && acc && labels && state && worklist, "null pointer");
    LEIDA_REQUIRE(metric == 0 || metric == 1, "metric: 0 cosine, 1 euclidean");
    int rc = check_runs(n_runs, N, x_pitch);
    if (rc) return rc;
    LEIDA_REQUIRE(kmax >= 2 && kmax <= LEIDA_KM_MAX_K, "2 <= kmax <= 64");
    LEIDA_REQUIRE(M >= 1 && M < (1LL << 31), "1 <= M < 2^31");
    LEIDA_REQUIRE(worklist_cap >= 0, "worklist_cap >= 0");
    AssignParams p;
    p.X = X; p.xnorm = xnorm; p.M = M; p.N = N; p.XP = (int)x_pitch;
    p.run_k = run_k; p.run_crow0 = run_crow0; p.cent = cent; p.cnorm = cnorm;
    p.acc = (long long*)acc; p.labels = labels; p.state = (long long*)state;
    p.worklist = worklist; p.worklist_cap = worklist_cap; p.labels_out = nullptr; p.metric = metric;
    // fp32 error bound of score = fl(dot)*fl(1/|c|): (n+2) roundings of relative size 2^-24 on
    // sum |x_j c_j| <= |x||c|, doubled for the difference of two scores, 10% slack.
    p.tau0 = 2.2f * (float)(x_pitch + 8) * 5.9604645e-8f;
    if (kmax <= 4) rc = launch_assign<4, 2>(p, n_runs, st);
    else if (kmax <= 8) rc = launch_assign<8, 2>(p, n_runs, st);
    else if (kmax <= 12) rc = launch_assign<12, 2>(p, n_runs, st);
    else if (kmax <= 16) rc = launch_assign<16, 2>(p, n_runs, st);
    else if (kmax <= 20) rc = launch_assign<20, 2>(p, n_runs, st);
    else if (kmax <= 24) rc = launch_assign<24, 2>(p, n_runs, st);
    else if (kmax <= 32) rc = launch_assign<32, 2>(p, n_runs, st);
    else if (kmax <= 40) rc = launch_assign<40, 2>(p, n_runs, st);
    else if (kmax <= 48) rc = launch_assign<48, 1>(p, n_runs, st);
    else rc = launch_assign<64, 1>(p, n_runs, st);
    return rc;
}

extern "C" int leida_kmeans_recheck(const float* X, const double* xnorm, int64_t M, int N, int64_t x_pitch,
                                    int n_runs, const int32_t* run_k, const int32_t* run_crow0, const float* cent,
                                    const double* cnorm, int64_t* acc, int32_t* labels, int64_t* state,
                                    int32_t* worklist, int64_t worklist_cap, int32_t* labels_out, int metric,
                                    leida_stream_t stream) {
    LEIDA_REQUIRE(X && xnorm && run_k && run_crow0 && cent && cnorm && acc && labels && state && worklist, "null pointer");
    LEIDA_REQUIRE(N >= 1 && x_pitch >= N && x_pitch % 32 == 0, "x_pitch multiple of 32 and >= N");
    LEIDA_REQUIRE(metric == 0 || metric == 1, "metric: 0 cosine, 1 euclidean");
    AssignParams p;
    p.X = X; p.xnorm = xnorm; p.M = M; p.N = N; p.XP = (int)x_pitch;
    p.run_k = run_k; p.run_crow0 = run_crow0; p.cent = cent; p.cnorm = cnorm;
    p.acc = (long long*)acc; p.labels = labels; p.state = (long long*)state;
    p.worklist = worklist; p.worklist_cap = worklist_cap; p.tau0 = 0.f; p.labels_out = labels_out; p.n_runs = n_runs;
    p.metric = metric;
    const int grid = sm_count() * (N > 256 ? 10 : 4);      // wide rows: more warps in flight hide the centroid loads
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 128) k_recheck<4><<<grid, 128, 0, st>>>(p);
    else if (N <= 256) k_recheck<8><<<grid, 128, 0, st>>>(p);
    else if (N <= 512) k_recheck<16><<<grid, 128, 0, st>>>(p);
    else k_recheck<0><<<grid, 128, 0, st>>>(p);
    return check_launch("k_recheck");
}

static int launch_update(int n_runs, const int32_t* run_k, const int32_t* run_crow0, int N, int64_t x_pitch,
                         const ReducedSource& red, float* cent, double* cnorm, int64_t* state, int n_iter,
                         int32_t* n_active_out, const int32_t* run_pcol, void* C1, void* C2, int32_t* mailbox,
                         leida_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_REQUIRE(run_k && run_crow0 && cent && cnorm && state && n_active_out, "null pointer");
    int rc = check_runs(n_runs, N, x_pitch);
    if (rc) return rc;
    LEIDA_REQUIRE(n_iter >= 0, "n_iter >= 0");
    LEIDA_REQUIRE((run_pcol == nullptr) == (C1 == nullptr) && (C1 == nullptr) == (C2 == nullptr), "run_pcol, C1, C2 together");
    const int NPK = (int)(((int64_t)N + 1 + 63) / 64 * 64);     // = leida_kmeans_tc_pitch(N)
    k_update<<<n_runs, 256, 0, st>>>(n_runs, run_k, run_crow0, N, (int)x_pitch, red, cent, cnorm, (long long*)state, n_iter,
                                     n_active_out, run_pcol, (__nv_bfloat16*)C1, (__nv_bfloat16*)C2, NPK, mailbox);
    return check_launch("k_update");
}

extern "C" int leida_kmeans_update(int n_runs, const int32_t* run_k, const int32_t* run_crow0, int N, int64_t x_pitch,
                                   const int64_t* acc_reduced, const int64_t* state_reduced, float* cent, double* cnorm,
                                   int64_t* state, int n_iter, int32_t* n_active_out, const int32_t* run_pcol, void* C1,
                                   void* C2, int32_t* mailbox, leida_stream_t stream) {
    LEIDA_REQUIRE(acc_reduced && state_reduced, "null pointer");
    ReducedSource red{(const long long*)acc_reduced, (const long long*)state_reduced, nullptr, 0, 0, nullptr, nullptr, nullptr};
    return launch_update(n_runs, run_k, run_crow0, N, x_pitch, red, cent, cnorm, state, n_iter, n_active_out, run_pcol,
                         C1, C2, mailbox, stream);
}

extern "C" int leida_kmeans_update_peers(int n_runs, const int32_t* run_k, const int32_t* run_crow0, int N,
                                         int64_t x_pitch, const int64_t* const* peer_bufs, int n_peers,
                                         const int64_t* multicast_buf, int64_t state_offset_words,
                                         const uint64_t* flags, const uint64_t* seq, float* cent, double* cnorm,
                                         int64_t* state, int n_iter, int32_t* n_active_out, const int32_t* run_pcol,
                                         void* C1, void* C2, int32_t* mailbox, leida_stream_t stream) {
    LEIDA_REQUIRE(peer_bufs && n_peers >= 1 && n_peers <= 64 && state_offset_words > 0, "peer buffers");
    LEIDA_REQUIRE((flags == nullptr) == (seq == nullptr), "flags and seq together");
    ReducedSource red{nullptr, nullptr, (const long long* const*)peer_bufs, n_peers, (long long)state_offset_words,
                      (const long long*)multicast_buf, (const unsigned long long*)flags, (const unsigned long long*)seq};
    return launch_update(n_runs, run_k, run_crow0, N, x_pitch, red, cent, cnorm, state, n_iter, n_active_out, run_pcol,
                         C1, C2, mailbox, stream);
}

extern "C" int leida_kmeans_exchange_publish(int n_runs, const int32_t* run_k, const int32_t* run_crow0, int64_t x_pitch,
                                             const int64_t* local_words, int64_t n_acc_words, int64_t* exchange_half,
                                             uint64_t* const* peer_flags, int n_peers, int my_rank, uint64_t* seq,
                                             int32_t* counter, leida_stream_t stream) {
    LEIDA_REQUIRE(run_k && run_crow0 && local_words && exchange_half && peer_flags && seq && counter, "null pointer");
    LEIDA_REQUIRE(n_runs >= 1 && n_runs <= 65535 && x_pitch % 32 == 0 && n_peers >= 1 && n_peers <= 64 &&
                      my_rank >= 0 && my_rank < n_peers && n_acc_words > 0, "sizes");
    k_exchange_publish<<<n_runs, 256, 0, (cudaStream_t)stream>>>(
        n_runs, run_k, run_crow0, (int)x_pitch + 8, (const long long*)local_words, (long long)n_acc_words,
        (long long*)exchange_half, (unsigned long long* const*)peer_flags, n_peers, my_rank, (unsigned long long*)seq, counter);
    return check_launch("k_exchange_publish");
}

extern "C" int leida_kmeans_update_fused(int n_runs, const int32_t* run_k, const int32_t* run_crow0, int N, int64_t x_pitch,
                                         int64_t* local_words, int64_t state_offset_words, int64_t* exchange_half,
                                         const int64_t* const* peer_halves, const int64_t* multicast_half,
                                         uint64_t* const* peer_flags, const uint64_t* flags_local, int64_t flag_pitch,
                                         int n_peers, int my_rank, uint64_t* seq, float* cent, double* cnorm,
                                         int64_t* state, int n_iter, int32_t* n_active_out, const int32_t* run_pcol,
                                         void* C1, void* C2, int32_t* mailbox, leida_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_REQUIRE(run_k && run_crow0 && local_words && cent && cnorm && state && n_active_out, "null pointer");
    int rc = check_runs(n_runs, N, x_pitch);
    if (rc) return rc;
    LEIDA_REQUIRE(n_iter >= 0 && state_offset_words > 0 && x_pitch <= 1024, "n_iter >= 0, x_pitch <= 1024");
    LEIDA_REQUIRE((run_pcol == nullptr) == (C1 == nullptr) && (C1 == nullptr) == (C2 == nullptr), "run_pcol, C1, C2 together");
    if (exchange_half) {
        LEIDA_REQUIRE(peer_halves && peer_flags && flags_local && seq && n_peers >= 1 && n_peers <= 64 && my_rank >= 0 &&
                          my_rank < n_peers && flag_pitch >= n_runs, "exchange arguments");
    }
    ExchangeParams x;
    x.my_half = (long long*)exchange_half; x.peers = (const long long* const*)peer_halves;
    x.mc = (const long long*)multicast_half; x.peer_flags = (unsigned long long* const*)peer_flags;
    x.flags_local = (const unsigned long long*)flags_local; x.seq = (unsigned long long*)seq;
    x.n_peers = n_peers; x.my_rank = my_rank; x.flag_pitch = (int)flag_pitch; x.state_off = (long long)state_offset_words;
    const int NPK = (int)(((int64_t)N + 1 + 63) / 64 * 64);     // = leida_kmeans_tc_pitch(N)
#define LEIDA_UF(CPLV)                                                                                                    \
    k_update_fused<CPLV><<<n_runs, 256, 0, st>>>(n_runs, run_k, run_crow0, N, (int)x_pitch, x, (long long*)local_words, cent, \
                                                 cnorm, (long long*)state, n_iter, n_active_out, run_pcol,               \
                                                 (__nv_bfloat16*)C1, (__nv_bfloat16*)C2, NPK, mailbox)
    if (x_pitch <= 128) LEIDA_UF(4);
    else if (x_pitch <= 256) LEIDA_UF(8);
    else if (x_pitch <= 416) LEIDA_UF(13);
    else if (x_pitch <= 512) LEIDA_UF(16);
    else LEIDA_UF(32);
#undef LEIDA_UF
    return check_launch("k_update_fused");
}

extern "C" int leida_kmeans_distortion(const float* X, const double* xnorm, int64_t M, int N, int64_t x_pitch, int n_runs,
                                       int kmax, const int32_t* run_k, const int32_t* run_crow0, const float* cent,
                                       const double* cnorm, const int32_t* labels, int64_t* state, int metric,
                                       leida_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_REQUIRE(X && xnorm && run_k && run_crow0 && cent && cnorm && labels && state, "null pointer");
    int rc = check_runs(n_runs, N, x_pitch);
    if (rc) return rc;
    LEIDA_REQUIRE(kmax <= LEIDA_KM_MAX_K, "kmax <= 64 (0: unknown, use the gather kernel)");
    k_zero_dist<<<(n_runs + 127) / 128, 128, 0, st>>>((long long*)state, n_runs);
    if (M > 0) {
        // tiled form when a run's centroids (staged in a ring) and the per-run sums fit in shared memory; kmax (the
        // largest K of the batch) sizes the ring and comes from the caller, so the call stays asynchronous
        // ring depth: as deep as shared memory allows, 2..4 runs
        const size_t per_stage = (size_t)kmax * x_pitch * sizeof(float) + (size_t)kmax * sizeof(double);
        const size_t fixed2 = (size_t)n_runs * (2 * sizeof(unsigned long long) + 2 * sizeof(int)) + 64;
        int nst = 4;
        while (nst > 2 && nst * per_stage + fixed2 > (size_t)max_smem_optin()) --nst;
        const size_t smem2 = nst * per_stage + fixed2;
        const bool tiled = metric == 0 && kmax >= 1 && x_pitch <= 512 && smem2 <= (size_t)max_smem_optin() &&
                           getenv("LEIDA_DISTORTION_LEGACY") == nullptr;
        // pipelined form (bulk copies + mbarriers): needs 16-byte aligned rows and CPL == ceil(x_pitch / 32)
        const char* form = getenv("LEIDA_DISTORTION_FORM");           // "piped" (default) | "ring" (the cp.async form)
        const bool want_piped = form == nullptr || form[0] != 'r';
        if (tiled && want_piped && x_pitch % 4 == 0 && ((uintptr_t)cent & 15) == 0) {
            const size_t stage_floats = ((size_t)kmax * x_pitch + 31) / 32 * 32;
            const size_t per_stage3 = stage_floats * sizeof(float) + (size_t)kmax * sizeof(double) + 2 * sizeof(uint64_t);
            const size_t fixed3 = fixed2 + 32 * sizeof(float) + 128;
            int nst3 = 6;
            while (nst3 > 2 && nst3 * per_stage3 + fixed3 > (size_t)max_smem_optin()) --nst3;
            const size_t smem3 = nst3 * per_stage3 + fixed3;
            if (smem3 <= (size_t)max_smem_optin()) {
#define LEIDA_D3(CPLV, QV)                                                                                                \
    case CPLV: {                                                                                                          \
        const long long n_tiles = (M + D2_WARPS * 4 * QV - 1) / (D2_WARPS * 4 * QV);                                      \
        const unsigned g3 = (unsigned)(n_tiles < sm_count() ? n_tiles : sm_count());                                      \
        LEIDA_CUDA_TRY(cudaFuncSetAttribute(k_distortion_piped<CPLV, QV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3)); \
        k_distortion_piped<CPLV, QV><<<g3, D2_THREADS, smem3, st>>>(X, xnorm, M, (int)x_pitch, n_runs, run_k, run_crow0, cent, \
                                                                    cnorm, labels, (long long*)state, kmax, nst3);        \
    } break
                switch ((int)((x_pitch + 31) / 32)) {
                    LEIDA_D3(1, 3); LEIDA_D3(2, 3); LEIDA_D3(3, 3); LEIDA_D3(4, 3);
                    LEIDA_D3(5, 2); LEIDA_D3(6, 2); LEIDA_D3(7, 2); LEIDA_D3(8, 2);
                    LEIDA_D3(9, 1); LEIDA_D3(10, 1); LEIDA_D3(11, 1); LEIDA_D3(12, 1);
                    LEIDA_D3(13, 1); LEIDA_D3(14, 1); LEIDA_D3(15, 1); LEIDA_D3(16, 1);
                    default: LEIDA_REQUIRE(false, "k_distortion_piped: x_pitch out of range");
                }
#undef LEIDA_D3
                return check_launch("k_distortion_piped", 2);
            }
        }
        if (tiled) {
#define LEIDA_D2(CPLV, QV, NSTV)                                                                                          \
    do {                                                                                                                  \
        const long long n_tiles = (M + D2_WARPS * 4 * QV - 1) / (D2_WARPS * 4 * QV);                                      \
        const unsigned g2 = (unsigned)(n_tiles < sm_count() ? n_tiles : sm_count());                                      \
        LEIDA_CUDA_TRY(cudaFuncSetAttribute(k_distortion_tiled<CPLV, QV, NSTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
        k_distortion_tiled<CPLV, QV, NSTV><<<g2, D2_THREADS, smem2, st>>>(X, xnorm, M, (int)x_pitch, n_runs, run_k, run_crow0, cent, \
                                                                          cnorm, labels, (long long*)state, kmax);       \
    } while (0)
#define LEIDA_D2N(CPLV, QV)                                                                                               \
    do {                                                                                                                  \
        if (nst == 4) LEIDA_D2(CPLV, QV, 4); else if (nst == 3) LEIDA_D2(CPLV, QV, 3); else LEIDA_D2(CPLV, QV, 2);        \
    } while (0)
            if (x_pitch <= 128) LEIDA_D2N(4, 3);             // narrow rows: three quads per warp (144 rows per tile)
            else if (x_pitch <= 256) LEIDA_D2N(8, 2);
            else if (x_pitch <= 416) LEIDA_D2N(13, 1);
            else LEIDA_D2N(16, 1);
#undef LEIDA_D2N
#undef LEIDA_D2
            return check_launch("k_distortion_tiled", 2);
        }
        const unsigned grid = (unsigned)((M + DIST_ROWS - 1) / DIST_ROWS);
        if (N <= 128) k_distortion<4><<<grid, 256, 0, st>>>(X, xnorm, M, N, (int)x_pitch, n_runs, run_crow0, cent, cnorm, labels, (long long*)state, metric);
        else if (N <= 256) k_distortion<8><<<grid, 256, 0, st>>>(X, xnorm, M, N, (int)x_pitch, n_runs, run_crow0, cent, cnorm, labels, (long long*)state, metric);
        else k_distortion<0><<<grid, 256, 0, st>>>   /* wider rows: occupancy beats register residency (measured at N=400) */(X, xnorm, M, N, (int)x_pitch, n_runs, run_crow0, cent, cnorm, labels, (long long*)state, metric);
    }
    return check_launch("k_distortion", M > 0 ? 2 : 1);
}

extern "C" int leida_kmeans_means_f32_sequential(const float* X, int64_t M, int N, int64_t x_pitch, int n_runs,
                                                 const int32_t* run_k, const int32_t* run_crow0, const int32_t* labels,
                                                 const int64_t* state, float* cent, double* cnorm,
                                                 leida_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_REQUIRE(X && run_k && run_crow0 && labels && state && cent && cnorm, "null pointer");
    int rc = check_runs(n_runs, N, x_pitch);
    if (rc) return rc;
    const int groups = (int)(x_pitch / 32);
    const long long warps = (long long)LEIDA_KM_MAX_K * groups;
    dim3 grid((unsigned)((warps * 32 + 127) / 128), (unsigned)n_runs);
    k_means_f32_seq<<<grid, 128, 0, st>>>(X, M, N, (int)x_pitch, run_k, run_crow0, labels, (const long long*)state, cent,
                                          groups, 0);
    rc = check_launch("k_means_f32_seq");
    if (rc) return rc;
    k_cent_norms<<<n_runs, 256, 0, st>>>(run_k, run_crow0, (const long long*)state, cent, N, (int)x_pitch, cnorm);
    return check_launch("k_cent_norms");
}

extern "C" int leida_kmeans_transform_f64(const float* X, int64_t M, int N, int64_t x_pitch, const float* cent, int K,
                                          int64_t c_pitch, double* dist, int32_t* labels, int metric,
                                          leida_stream_t stream) {
    LEIDA_REQUIRE(X && cent, "null pointer");
    LEIDA_REQUIRE(metric == 0 || metric == 1, "metric: 0 cosine, 1 euclidean");
    LEIDA_REQUIRE(dist || labels, "at least one output");
    LEIDA_REQUIRE(M >= 0 && N >= 1 && K >= 1 && x_pitch >= N && c_pitch >= N, "sizes");
    if (M == 0) return LEIDA_OK;
    const long long threads = M * 32;
    k_transform<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, M, N, x_pitch, cent, K, c_pitch,
                                                                                     dist, labels, metric);
    return check_launch("k_transform");
}

extern "C" int leida_remap_labels_i32(int32_t* labels, int64_t M, const int32_t* lut, int K, leida_stream_t stream) {
    LEIDA_REQUIRE(labels && lut, "null pointer");
    LEIDA_REQUIRE(M >= 0 && K >= 1, "sizes");
    if (M == 0) return LEIDA_OK;
    k_remap<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(labels, M, lut, K);
    return check_launch("k_remap");
}

extern "C" int leida_remap_labels_batch_i32(const int32_t* labels, int64_t M, int n_out, const int32_t* src_run,
                                            const int32_t* lut, int kmax, int32_t* out, leida_stream_t stream) {
    LEIDA_REQUIRE(labels && src_run && lut && out, "null pointer");
    LEIDA_REQUIRE(M >= 0 && n_out >= 1 && n_out <= 65535 && kmax >= 1, "sizes");
    if (M == 0) return LEIDA_OK;
    dim3 grid((unsigned)((M + 255) / 256), (unsigned)n_out);
    k_remap_batch<<<grid, 256, 0, (cudaStream_t)stream>>>(labels, M, src_run, lut, kmax, out);
    return check_launch("k_remap_batch");
}
