// General path of stage 2b: get_eigenvectors(dFC, n=1) for an arbitrary symmetric (N,N,T') tensor
// (pyleida/signal_tools/_signal_tools.py:185-224; eigs(A, 1, which='LM') at :213).
//
// dFC is laid out with the volume index fastest, so one *thread per volume* reads it perfectly
// coalesced and every volume's eigenproblem advances in lock-step:
//     Y = A Q            one streaming pass over dFC per iteration (the only HBM-sized traffic)
//     Rayleigh-Ritz      B = Q^T Y (4x4, Jacobi), Ritz pair of largest |lambda|, residual test
//     Q = orth(Y)        modified Gram-Schmidt
// Block size P = 4: the convergence factor is |lambda_5/lambda_1|, and a rank-2 phase-coherence
// matrix is solved exactly by the second pass.  All per-volume state is stored volume-fastest.
#include "common.cuh"
#include <math.h>

namespace leida {

constexpr int P = 4;     // subspace width
constexpr int IB = 8;    // rows of A per thread in the matmul

__device__ __forceinline__ double hash_unit(unsigned a, unsigned b) {
    unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return ((double)h + 0.5) * (1.0 / 4294967296.0) - 0.5;
}

// Y[j][c][t] = deterministic pseudo-random start block (same for every volume).
__global__ void k_dense_seed(double* Y, int N, int Tp) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tp) return;
    for (int j = 0; j < N; ++j)
        for (int c = 0; c < P; ++c) Y[((long long)j * P + c) * Tp + t] = hash_unit(j, c);
}

// Q = orth(Y) per volume (modified Gram-Schmidt, columns that vanish are zeroed).
__global__ void k_dense_ortho(const double* __restrict__ Y, double* __restrict__ Q, int N, int Tp,
                              const int32_t* __restrict__ info, int gate) {
    if (gate && info[1] == 0) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tp) return;
    double scale = 0.0;
    for (int c = 0; c < P; ++c) {
        // q_c = y_c - sum_{d<c} (q_d . y_c) q_d   (two passes: project, then normalise)
        double proj[P];
        for (int d = 0; d < c; ++d) proj[d] = 0.0;
        for (int j = 0; j < N; ++j) {
            const double y = Y[((long long)j * P + c) * Tp + t];
            for (int d = 0; d < c; ++d) proj[d] += Q[((long long)j * P + d) * Tp + t] * y;
        }
        double nn = 0.0;
        for (int j = 0; j < N; ++j) {
            double y = Y[((long long)j * P + c) * Tp + t];
            for (int d = 0; d < c; ++d) y -= proj[d] * Q[((long long)j * P + d) * Tp + t];
            Q[((long long)j * P + c) * Tp + t] = y;
            nn += y * y;
        }
        // re-orthogonalise once (classical "twice is enough")
        for (int d = 0; d < c; ++d) proj[d] = 0.0;
        for (int j = 0; j < N; ++j) {
            const double y = Q[((long long)j * P + c) * Tp + t];
            for (int d = 0; d < c; ++d) proj[d] += Q[((long long)j * P + d) * Tp + t] * y;
        }
        nn = 0.0;
        for (int j = 0; j < N; ++j) {
            double y = Q[((long long)j * P + c) * Tp + t];
            for (int d = 0; d < c; ++d) y -= proj[d] * Q[((long long)j * P + d) * Tp + t];
            Q[((long long)j * P + c) * Tp + t] = y;
            nn += y * y;
        }
        const double nrm = sqrt(nn);
        if (c == 0) scale = nrm;
        const double inv = (nrm > 1e-13 * scale && nrm > 0.0) ? 1.0 / nrm : 0.0;
        for (int j = 0; j < N; ++j) Q[((long long)j * P + c) * Tp + t] *= inv;
    }
}

// Y = A Q.  grid (ceil(N/IB), ceil(Tp/128)); one thread per volume, IB rows of A per thread.
__global__ void __launch_bounds__(128)
k_dense_matmul(const double* __restrict__ A, const double* __restrict__ Q, double* __restrict__ Y, int N, int Tp,
               const int32_t* __restrict__ info, int gate) {
    if (gate && info[1] == 0) return;
    const int t = blockIdx.y * blockDim.x + threadIdx.x;
    if (t >= Tp) return;
    const int i0 = blockIdx.x * IB;
    double acc[IB][P];
#pragma unroll
    for (int a = 0; a < IB; ++a)
#pragma unroll
        for (int c = 0; c < P; ++c) acc[a][c] = 0.0;
    for (int j = 0; j < N; ++j) {
        double q[P];
#pragma unroll
        for (int c = 0; c < P; ++c) q[c] = Q[((long long)j * P + c) * Tp + t];
#pragma unroll
        for (int a = 0; a < IB; ++a) {
            const int i = i0 + a;
            const double v = i < N ? A[((long long)i * N + j) * Tp + t] : 0.0;
#pragma unroll
            for (int c = 0; c < P; ++c) acc[a][c] += v * q[c];
        }
    }
#pragma unroll
    for (int a = 0; a < IB; ++a)
        if (i0 + a < N)
#pragma unroll
            for (int c = 0; c < P; ++c) Y[((long long)(i0 + a) * P + c) * Tp + t] = acc[a][c];
}

// Rayleigh-Ritz per volume: B = Q^T Y, Jacobi, largest-|lambda| Ritz vector -> V[j][t], residual test.
__global__ void k_dense_ritz(const double* __restrict__ Q, const double* __restrict__ Y, double* __restrict__ V,
                             int N, int Tp, double tol, int32_t* info, int gate, int iter) {
    if (gate && info[1] == 0 && iter > 0) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tp) return;
    double B[P][P];
    for (int a = 0; a < P; ++a)
        for (int b = 0; b < P; ++b) B[a][b] = 0.0;
    for (int j = 0; j < N; ++j) {
        double q[P], y[P];
        for (int c = 0; c < P; ++c) {
            q[c] = Q[((long long)j * P + c) * Tp + t];
            y[c] = Y[((long long)j * P + c) * Tp + t];
        }
        for (int a = 0; a < P; ++a)
            for (int b = 0; b < P; ++b) B[a][b] += q[a] * y[b];
    }
    for (int a = 0; a < P; ++a)
        for (int b = a + 1; b < P; ++b) B[a][b] = B[b][a] = 0.5 * (B[a][b] + B[b][a]);
    double S[P][P];
    for (int a = 0; a < P; ++a)
        for (int b = 0; b < P; ++b) S[a][b] = a == b ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0.0;
        for (int a = 0; a < P; ++a)
            for (int b = a + 1; b < P; ++b) off += B[a][b] * B[a][b];
        if (off < 1e-300) break;
        for (int a = 0; a < P; ++a)
            for (int b = a + 1; b < P; ++b) {
                if (B[a][b] == 0.0) continue;
                const double theta = (B[b][b] - B[a][a]) / (2.0 * B[a][b]);
                const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double cs = 1.0 / sqrt(tt * tt + 1.0), sn = tt * cs;
                for (int k = 0; k < P; ++k) {
                    const double bka = B[k][a], bkb = B[k][b];
                    B[k][a] = cs * bka - sn * bkb;
                    B[k][b] = sn * bka + cs * bkb;
                }
                for (int k = 0; k < P; ++k) {
                    const double bak = B[a][k], bbk = B[b][k];
                    B[a][k] = cs * bak - sn * bbk;
                    B[b][k] = sn * bak + cs * bbk;
                }
                for (int k = 0; k < P; ++k) {
                    const double ska = S[k][a], skb = S[k][b];
                    S[k][a] = cs * ska - sn * skb;
                    S[k][b] = sn * ska + cs * skb;
                }
            }
    }
    int best = 0;
    for (int a = 1; a < P; ++a)
        if (fabs(B[a][a]) > fabs(B[best][best])) best = a;
    const double lam = B[best][best];
    double s[P];
    for (int c = 0; c < P; ++c) s[c] = S[c][best];
    double r2 = 0.0, v2 = 0.0;
    for (int j = 0; j < N; ++j) {
        double v = 0.0, av = 0.0;
        for (int c = 0; c < P; ++c) {
            v += Q[((long long)j * P + c) * Tp + t] * s[c];
            av += Y[((long long)j * P + c) * Tp + t] * s[c];
        }
        V[(long long)j * Tp + t] = v;
        const double r = av - lam * v;
        r2 += r * r;
        v2 += v * v;
    }
    const bool ok = (v2 > 0.0) && (sqrt(r2) <= tol * fabs(lam) * sqrt(v2));
    if (!ok) atomicAdd(&info[2], 1);
}

__global__ void k_dense_roll_info(int32_t* info, int iter) {
    // info[1] = unconverged volumes after this iteration; info[0] = iterations that did work
    if (info[1] != 0 || iter == 0) info[0] = iter + 1;
    info[1] = info[2];
    info[2] = 0;
}

// Sign rule + normalisation + transpose to (T', N) row-major.  One CTA per 32 volumes.
__global__ void __launch_bounds__(256)
k_dense_finalize(const double* __restrict__ V, int N, int Tp, double* __restrict__ out, long long out_pitch) {
    __shared__ double tile[32][33];
    __shared__ double s_scale[32];
    const int t0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;   // 8 warps
    // statistics: warp w handles volumes t0 + w*4 .. +3 ; lanes stride over j -- V[j][t] is volume-fastest,
    // so instead let lane = volume and loop j in the warp-strided fashion.
    if (w == 0) {
        const int t = t0 + lane;
        double np = 0.0, sp = 0.0, sn = 0.0, nn = 0.0;
        if (t < Tp)
            for (int j = 0; j < N; ++j) {
                const double v = V[(long long)j * Tp + t];
                nn += v * v;
                if (v > 0.0) { np += 1.0; sp += v; }
                else if (v < 0.0) sn += v;
            }
        double sc = nn > 0.0 ? 1.0 / sqrt(nn) : 0.0;
        if (2.0 * np > (double)N) sc = -sc;
        else if (2.0 * np == (double)N && sp > -sn) sc = -sc;
        s_scale[lane] = sc;
    }
    __syncthreads();
    for (int j0 = 0; j0 < N; j0 += 32) {
        for (int r = w; r < 32; r += 8) {
            const int j = j0 + r, t = t0 + lane;
            tile[r][lane] = (j < N && t < Tp) ? V[(long long)j * Tp + t] : 0.0;
        }
        __syncthreads();
        for (int r = w; r < 32; r += 8) {
            const int t = t0 + r, j = j0 + lane;
            if (t < Tp && j < N) out[(long long)t * out_pitch + j] = tile[lane][r] * s_scale[r];
        }
        __syncthreads();
    }
}

}  // namespace leida

using namespace leida;

extern "C" size_t leida_eigvec_dense_workspace_bytes(int N, int n_vol) {
    if (N < 1 || n_vol < 1) return 0;
    return ((size_t)2 * N * P + (size_t)N) * (size_t)n_vol * sizeof(double);
}

extern "C" int leida_eigvec_dense_f64(const double* dfc, int N, int n_vol, double* out, int64_t out_pitch, double tol,
                                      int max_iter, int32_t* info, void* workspace, size_t workspace_bytes,
                                      leida_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_REQUIRE(dfc && out && info && workspace, "null pointer");
    LEIDA_REQUIRE(N >= 1 && n_vol >= 0 && out_pitch >= N, "sizes");
    LEIDA_REQUIRE(tol > 0.0 && max_iter >= 1, "tol > 0 and max_iter >= 1");
    if (workspace_bytes < leida_eigvec_dense_workspace_bytes(N, n_vol)) {
        set_error("dense eigenvector workspace too small");
        return LEIDA_ERR_WORKSPACE;
    }
    if (n_vol == 0) return LEIDA_OK;
    const int Tp = n_vol;
    double* Q = (double*)workspace;
    double* Y = Q + (size_t)N * P * Tp;
    double* V = Y + (size_t)N * P * Tp;
    // info: [0] iterations, [1] unconverged, [2] scratch counter -- the caller provides >= 4 ints
    LEIDA_CUDA_TRY(cudaMemsetAsync(info, 0, 4 * sizeof(int32_t), st));
    const int tb = 128, tg = (Tp + tb - 1) / tb;
    k_dense_seed<<<tg, tb, 0, st>>>(Y, N, Tp);
    k_dense_ortho<<<tg, tb, 0, st>>>(Y, Q, N, Tp, info, 0);
    dim3 mg((N + IB - 1) / IB, tg);
    for (int it = 0; it < max_iter; ++it) {
        const int gate = it > 0;
        k_dense_matmul<<<mg, tb, 0, st>>>(dfc, Q, Y, N, Tp, info, gate);
        k_dense_ritz<<<tg, tb, 0, st>>>(Q, Y, V, N, Tp, tol, info, gate, it);
        k_dense_ortho<<<tg, tb, 0, st>>>(Y, Q, N, Tp, info, gate);
        k_dense_roll_info<<<1, 1, 0, st>>>(info, it);
    }
    k_dense_finalize<<<(Tp + 31) / 32, 256, 0, st>>>(V, N, Tp, out, out_pitch);
    return check_launch("leida_eigvec_dense_f64", 3 + 4 * max_iter);
}
