// Stage 1 + 2 of the LEiDA hot path on sm_100a:
//   * batched Hilbert phase (shared-memory mixed-radix / Bluestein FFT in fp64)
//       replaces pyleida/signal_tools/_signal_tools.py:24-44 (scipy.signal.hilbert + np.angle)
//   * phase coherence (standalone writer)          _signal_tools.py:137-183
//   * rank-2 closed-form leading eigenvector       _signal_tools.py:137-224 (dFC never materialised)
#include "common.cuh"
#include <math.h>

namespace leida {

// ------------------------------------------------------------------------------------------------
// FFT machinery.  Stockham autosort, decimation in frequency, one CTA per transform, data in shared
// memory as double2.  Twiddles tw[k] = exp(-2 pi i k / n) live in global memory (L1-resident).
// ------------------------------------------------------------------------------------------------
struct FftPlan {
    int n;
    int nfac;
    int radix[20];
};

static bool make_plan(int n, FftPlan& p) {
    p.n = n;
    p.nfac = 0;
    int m = n;
    while (m % 4 == 0) { p.radix[p.nfac++] = 4; m /= 4; }
    while (m % 2 == 0) { p.radix[p.nfac++] = 2; m /= 2; }
    while (m % 3 == 0) { p.radix[p.nfac++] = 3; m /= 3; }
    while (m % 5 == 0) { p.radix[p.nfac++] = 5; m /= 5; }
    for (int f = 7; f <= 31 && m > 1; f += 2)
        while (m % f == 0) { p.radix[p.nfac++] = f; m /= f; }
    return m == 1;
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// m - i*n and m + i*n
__device__ __forceinline__ double2 c_mi(double2 m, double2 n) { return make_double2(m.x + n.y, m.y - n.x); }
__device__ __forceinline__ double2 c_pi(double2 m, double2 n) { return make_double2(m.x - n.y, m.y + n.x); }

template <int R>
__device__ __forceinline__ void butterfly(double2* v);

template <>
__device__ __forceinline__ void butterfly<2>(double2* v) {
    double2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void butterfly<4>(double2* v) {
    double2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    double2 t2 = cadd(v[1], v[3]), t3 = csub(v[1], v[3]);
    v[0] = cadd(t0, t2);
    v[2] = csub(t0, t2);
    v[1] = c_mi(t1, t3);
    v[3] = c_pi(t1, t3);
}
template <>
__device__ __forceinline__ void butterfly<3>(double2* v) {
    const double h = 0.86602540378443864676;  // sqrt(3)/2
    double2 t1 = cadd(v[1], v[2]);
    double2 t2 = make_double2(v[0].x - 0.5 * t1.x, v[0].y - 0.5 * t1.y);
    double2 d = csub(v[1], v[2]);
    double2 t3 = make_double2(h * d.x, h * d.y);
    v[0] = cadd(v[0], t1);
    v[1] = c_mi(t2, t3);
    v[2] = c_pi(t2, t3);
}
template <>
__device__ __forceinline__ void butterfly<5>(double2* v) {
    const double c1 = 0.30901699437494742410;   // cos(2pi/5)
    const double c2 = -0.80901699437494742410;  // cos(4pi/5)
    const double s1 = 0.95105651629515357212;   // sin(2pi/5)
    const double s2 = 0.58778525229247312917;   // sin(4pi/5)
    double2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    double2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    double2 m1 = make_double2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    double2 m2 = make_double2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    double2 n1 = make_double2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
    double2 n2 = make_double2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
    v[0] = make_double2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
    v[1] = c_mi(m1, n1);
    v[4] = c_pi(m1, n1);
    v[2] = c_mi(m2, n2);
    v[3] = c_pi(m2, n2);
}

// One Stockham pass of radix R: n/R butterflies spread over the CTA.
template <int R>
__device__ __forceinline__ void fft_pass(const double2* __restrict__ in, double2* __restrict__ out,
                                         const double2* __restrict__ tw, int n, int len, int s) {
    const int m = len / R, nb = n / R, tstep = n / len;
    for (int idx = threadIdx.x; idx < nb; idx += blockDim.x) {
        const int p = idx / s, q = idx - p * s;
        double2 v[R];
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = in[q + s * (p + k * m)];
        butterfly<R>(v);
        double2* o = out + q + s * (R * p);
        o[0] = v[0];
#pragma unroll
        for (int j = 1; j < R; ++j) o[s * j] = cmul(v[j], __ldg(&tw[p * j * tstep]));
    }
}

// Generic odd prime radix (7..31): O(r^2) butterfly from the twiddle table.
__device__ __noinline__ void fft_pass_generic(const double2* __restrict__ in, double2* __restrict__ out,
                                              const double2* __restrict__ tw, int n, int len, int s, int r) {
    const int m = len / r, nb = n / r, tstep = n / len, rstep = n / r;
    for (int idx = threadIdx.x; idx < nb; idx += blockDim.x) {
        const int p = idx / s, q = idx - p * s;
        double2 v[31];
        for (int k = 0; k < r; ++k) v[k] = in[q + s * (p + k * m)];
        for (int j = 0; j < r; ++j) {
            double2 acc = v[0];
            for (int k = 1; k < r; ++k) acc = cadd(acc, cmul(v[k], __ldg(&tw[((j * k) % r) * rstep])));
            out[q + s * (r * p + j)] = cmul(acc, __ldg(&tw[p * j * tstep]));
        }
    }
}

// Forward DFT of plan.n points held in `src`; on return `src` points at the result.
__device__ __forceinline__ void fft_forward(double2*& src, double2*& dst, const double2* __restrict__ tw,
                                            const FftPlan& plan) {
    int len = plan.n, s = 1;
    for (int f = 0; f < plan.nfac; ++f) {
        const int r = plan.radix[f];
        switch (r) {
            case 4: fft_pass<4>(src, dst, tw, plan.n, len, s); break;
            case 2: fft_pass<2>(src, dst, tw, plan.n, len, s); break;
            case 3: fft_pass<3>(src, dst, tw, plan.n, len, s); break;
            case 5: fft_pass<5>(src, dst, tw, plan.n, len, s); break;
            default: fft_pass_generic(src, dst, tw, plan.n, len, s, r); break;
        }
        __syncthreads();
        double2* t = src; src = dst; dst = t;
        len /= r;
        s *= r;
    }
}

// Bluestein tables: w[n] = exp(-i pi n^2 / T) (T entries), Bhat = FFT_L(b), b[n] = conj(w[|n|]).
struct BluesteinTables {
    const double2* chirp;
    const double2* bhat;
    int L;
};

// Forward DFT of T points (arbitrary T) held in src[0..T): src/dst are buffers of >= L points.
__device__ __forceinline__ void dft_bluestein(double2*& src, double2*& dst, int T, const double2* __restrict__ twL,
                                              const FftPlan& planL, const BluesteinTables& bt) {
    const int L = bt.L;
    for (int n = threadIdx.x; n < L; n += blockDim.x)
        src[n] = n < T ? cmul(src[n], __ldg(&bt.chirp[n])) : make_double2(0.0, 0.0);
    __syncthreads();
    fft_forward(src, dst, twL, planL);
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
        double2 c = cmul(src[k], __ldg(&bt.bhat[k]));
        src[k] = make_double2(c.x, -c.y);
    }
    __syncthreads();
    fft_forward(src, dst, twL, planL);
    const double invL = 1.0 / (double)L;
    for (int k = threadIdx.x; k < T; k += blockDim.x) {
        double2 c = make_double2(src[k].x * invL, -src[k].y * invL);
        src[k] = cmul(c, __ldg(&bt.chirp[k]));
    }
    __syncthreads();
}

__global__ void k_twiddles(double2* tw, int n) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        double s, c;
        sincospi(2.0 * (double)k / (double)n, &s, &c);
        tw[k] = make_double2(c, -s);
    }
}

__global__ void k_chirp(double2* w, int T) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < T) {
        long long r = ((long long)n * n) % (2LL * T);
        double s, c;
        sincospi((double)r / (double)T, &s, &c);
        w[n] = make_double2(c, -s);
    }
}

// One CTA: bhat = FFT_L(b_pad).
__global__ void k_bluestein_bhat(const double2* __restrict__ chirp, const double2* __restrict__ twL, double2* bhat,
                                 int T, FftPlan planL) {
    extern __shared__ double2 sm[];
    const int L = planL.n;
    double2* a = sm;
    double2* b = sm + L;
    for (int n = threadIdx.x; n < L; n += blockDim.x) {
        double2 v = make_double2(0.0, 0.0);
        if (n < T) {
            double2 w = chirp[n];
            v = make_double2(w.x, -w.y);
        } else if (L - n < T) {
            double2 w = chirp[L - n];
            v = make_double2(w.x, -w.y);
        }
        a[n] = v;
    }
    __syncthreads();
    fft_forward(a, b, twL, planL);
    for (int n = threadIdx.x; n < L; n += blockDim.x) bhat[n] = a[n];
}

// Hilbert phase of two real series per transform: z = x1 + i x2; A(z) = ifft(h fft(z)) =
// (x1 - H x2) + i (x2 + H x1)  =>  H x1 = Im A - x2,  H x2 = x1 - Re A;  theta = atan2(H x, x).
template <bool BLUESTEIN>
__global__ void k_hilbert_phase(const double* __restrict__ x, long long n_series, int T, long long x_pitch,
                                double* __restrict__ phase, long long ph_pitch, int t_begin, int t_count,
                                const double2* __restrict__ tw, FftPlan plan, BluesteinTables bt) {
    extern __shared__ double2 sm[];
    const int bufn = BLUESTEIN ? bt.L : T;
    const long long n_pairs = (n_series + 1) / 2;
    const int half = (T + 1) / 2;
    const double invT = 1.0 / (double)T;
    for (long long pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        double2* src = sm;
        double2* dst = sm + bufn;
        const double* x1 = x + (2 * pair) * x_pitch;
        const bool has2 = (2 * pair + 1) < n_series;
        const double* x2 = x1 + (has2 ? x_pitch : 0);
        for (int t = threadIdx.x; t < T; t += blockDim.x) src[t] = make_double2(x1[t], has2 ? x2[t] : 0.0);
        __syncthreads();
        if (BLUESTEIN) dft_bluestein(src, dst, T, tw, plan, bt);
        else fft_forward(src, dst, tw, plan);
        for (int k = threadIdx.x; k < T; k += blockDim.x) {
            double h = (k == 0 || (2 * k == T)) ? 1.0 : (k < half ? 2.0 : 0.0);
            double2 v = src[k];
            src[k] = make_double2(h * v.x, -h * v.y);
        }
        __syncthreads();
        if (BLUESTEIN) dft_bluestein(src, dst, T, tw, plan, bt);
        else fft_forward(src, dst, tw, plan);
        double* o1 = phase + (2 * pair) * ph_pitch;
        double* o2 = o1 + ph_pitch;
        for (int i = threadIdx.x; i < t_count; i += blockDim.x) {
            const int t = t_begin + i;
            const double2 r = src[t];
            const double are = r.x * invT, aim = -r.y * invT;
            const double v1 = x1[t], v2 = has2 ? x2[t] : 0.0;
            o1[i] = atan2(aim - v2, v1);
            if (has2) o2[i] = atan2(v1 - are, v2);
        }
        __syncthreads();
    }
}

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ------------------------------------------------------------------------------------------------
// phase coherence writer: dFC[i][j][t] = cos(th_i - th_j) = c_i c_j + s_i s_j, t fastest.
// ------------------------------------------------------------------------------------------------
__global__ void k_phase_coherence(const double* __restrict__ phase, int N, int T, long long pitch,
                                  double* __restrict__ dfc) {
    const int Tp = T - 2;
    const int i = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tp) return;
    double si, ci;
    sincos(phase[(long long)i * pitch + t + 1], &si, &ci);
    double* out = dfc + ((long long)i * N) * Tp + t;
    for (int j = 0; j < N; ++j) {
        double sj, cj;
        sincos(phase[(long long)j * pitch + t + 1], &sj, &cj);
        out[(long long)j * Tp] = ci * cj + si * sj;
    }
}

// ------------------------------------------------------------------------------------------------
// rank-2 leading eigenvector.  One CTA per (subject, tile of TT volumes); cos/sin of the tile kept
// in shared memory ([TT][NPAD] doubles each, NPAD odd to keep the transposed stores conflict-light).
//   z_t = sum_j exp(2 i th_j);  phi = arg(z)/2;  v_j = cos(th_j - phi) / sqrt(N/2 + |z|/2)
//   sign rule of _signal_tools.py:215-220.
// ------------------------------------------------------------------------------------------------
template <int TT>
__global__ void __launch_bounds__(256)
k_eigvec_rank2(const double* __restrict__ phase, int N, int n_vol, long long subj_stride, long long row_pitch,
               double* __restrict__ lei, long long lei_pitch, float* __restrict__ xf, long long xf_pitch) {
    extern __shared__ double smd[];
    const int NPAD = N | 1;
    double* sc = smd;                       // [TT][NPAD]
    double* ss = smd + (size_t)TT * NPAD;   // [TT][NPAD]
    double* red = ss + (size_t)TT * NPAD;   // [4][256] partials
    double* par = red + 4 * 256;            // [TT][4]: cos phi, sin phi, scale(with sign), unused
    const int tid = threadIdx.x;
    const int tt = tid % TT;
    const int v0 = blockIdx.x * TT;
    const long long s = blockIdx.y;
    const double* ph = phase + s * subj_stride;
    const bool live = (v0 + tt) < n_vol;

    double c2 = 0.0, s2 = 0.0;
    for (int e = tid; e < N * TT; e += 256) {
        const int j = e / TT;
        double sn = 0.0, cs = 0.0;
        if (live) sincos(ph[(long long)j * row_pitch + v0 + tt], &sn, &cs);
        sc[tt * NPAD + j] = cs;
        ss[tt * NPAD + j] = sn;
        c2 += cs * cs - sn * sn;
        s2 += 2.0 * cs * sn;
    }
    red[tid] = c2;
    red[256 + tid] = s2;
    __syncthreads();
    if (tid < TT) {
        double a = 0.0, b = 0.0;
        for (int i = tid; i < 256; i += TT) { a += red[i]; b += red[256 + i]; }
        const double r = hypot(a, b);
        const double phi = 0.5 * atan2(b, a);
        double sp, cp;
        sincos(phi, &sp, &cp);
        par[tid * 4 + 0] = cp;
        par[tid * 4 + 1] = sp;
        par[tid * 4 + 2] = rsqrt(0.5 * (double)N + 0.5 * r);
    }
    __syncthreads();
    // sign rule statistics
    {
        const double cp = par[tt * 4 + 0], sp = par[tt * 4 + 1];
        double npos = 0.0, spos = 0.0, sneg = 0.0;
        for (int e = tid; e < N * TT; e += 256) {
            const int j = e / TT;
            const double v = sc[tt * NPAD + j] * cp + ss[tt * NPAD + j] * sp;
            if (v > 0.0) { npos += 1.0; spos += v; }
            else if (v < 0.0) sneg += v;
        }
        __syncthreads();
        red[tid] = npos;
        red[256 + tid] = spos;
        red[512 + tid] = sneg;
    }
    __syncthreads();
    if (tid < TT) {
        double np = 0.0, sp_ = 0.0, sn_ = 0.0;
        for (int i = tid; i < 256; i += TT) { np += red[i]; sp_ += red[256 + i]; sn_ += red[512 + i]; }
        bool flip = false;
        if (2.0 * np > (double)N) flip = true;
        else if (2.0 * np == (double)N && sp_ > -sn_) flip = true;
        if (flip) par[tid * 4 + 2] = -par[tid * 4 + 2];
    }
    __syncthreads();
    const long long row0 = s * (long long)n_vol + v0;
    const int XP = (int)xf_pitch;
    if (lei) {
        for (int e = tid; e < N * TT; e += 256) {
            const int t = e / N, j = e - t * N;
            if (v0 + t < n_vol) {
                const double v = (sc[t * NPAD + j] * par[t * 4 + 0] + ss[t * NPAD + j] * par[t * 4 + 1]) * par[t * 4 + 2];
                lei[(row0 + t) * lei_pitch + j] = v;
            }
        }
    }
    if (xf) {
        for (int e = tid; e < XP * TT; e += 256) {
            const int t = e / XP, j = e - t * XP;
            if (v0 + t < n_vol) {
                float v = 0.f;
                if (j < N)
                    v = (float)((sc[t * NPAD + j] * par[t * 4 + 0] + ss[t * NPAD + j] * par[t * 4 + 1]) * par[t * 4 + 2]);
                xf[(row0 + t) * xf_pitch + j] = v;
            }
        }
    }
}

template <int TT>
static int launch_eigvec(const double* phase, int64_t S, int N, int n_vol, int64_t subj_stride, int64_t row_pitch,
                         double* lei, int64_t lei_pitch, float* xf, int64_t xf_pitch, cudaStream_t st) {
    const size_t smem = ((size_t)2 * TT * (N | 1) + 4 * 256 + 4 * TT) * sizeof(double);
    LEIDA_CUDA_TRY(cudaFuncSetAttribute(k_eigvec_rank2<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((n_vol + TT - 1) / TT, (unsigned)S);
    k_eigvec_rank2<TT><<<grid, 256, smem, st>>>(phase, N, n_vol, subj_stride, row_pitch, lei, lei_pitch, xf, xf_pitch);
    return check_launch("k_eigvec_rank2");
}

}  // namespace leida

using namespace leida;

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" size_t leida_hilbert_workspace_bytes(int T) {
    if (T < 1) return 0;
    FftPlan p;
    if (make_plan(T, p)) return (size_t)T * sizeof(double2);
    const int L = next_pow2(2 * T - 1);
    return ((size_t)L * 2 + (size_t)T) * sizeof(double2);
}

extern "C" int leida_hilbert_phase_f64(const double* x, int64_t n_series, int T, int64_t x_pitch, double* phase,
                                       int64_t phase_pitch, int t_begin, int t_count, void* workspace,
                                       size_t workspace_bytes, leida_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_REQUIRE(x && phase && workspace, "null pointer");
    LEIDA_REQUIRE(n_series >= 0 && T >= 1, "n_series >= 0 and T >= 1");
    LEIDA_REQUIRE(x_pitch >= T, "x_pitch >= T");
    LEIDA_REQUIRE(t_begin >= 0 && t_count >= 0 && t_begin + t_count <= T, "volume window inside [0,T)");
    LEIDA_REQUIRE(phase_pitch >= t_count, "phase_pitch >= t_count");
    if (workspace_bytes < leida_hilbert_workspace_bytes(T)) {
        set_error("hilbert workspace too small: %zu < %zu", workspace_bytes, leida_hilbert_workspace_bytes(T));
        return LEIDA_ERR_WORKSPACE;
    }
    if (n_series == 0 || t_count == 0) return LEIDA_OK;
    FftPlan plan;
    const bool direct = make_plan(T, plan);
    const int smem_max = max_smem_optin();
    const long long n_pairs = (n_series + 1) / 2;
    BluesteinTables bt{nullptr, nullptr, 0};
    double2* tw = (double2*)workspace;
    size_t smem;
    int threads;
    if (direct) {
        smem = (size_t)2 * T * sizeof(double2);
        if ((long long)smem > smem_max) {
            set_error("T=%d too long for the shared-memory FFT (max %d)", T, smem_max / 32);
            return LEIDA_ERR_UNSUPPORTED;
        }
        k_twiddles<<<(T + 255) / 256, 256, 0, st>>>(tw, T);
        threads = T >= 1024 ? 256 : (T >= 256 ? 128 : 64);
    } else {
        const int L = next_pow2(2 * T - 1);
        smem = (size_t)2 * L * sizeof(double2);
        if ((long long)smem > smem_max) {
            set_error("T=%d has a prime factor > 31 and is too long for the Bluestein path (max 2048)", T);
            return LEIDA_ERR_UNSUPPORTED;
        }
        FftPlan planL;
        make_plan(L, planL);
        double2* chirp = tw + L;
        double2* bhat = chirp + T;
        k_twiddles<<<(L + 255) / 256, 256, 0, st>>>(tw, L);
        k_chirp<<<(T + 255) / 256, 256, 0, st>>>(chirp, T);
        LEIDA_CUDA_TRY(cudaFuncSetAttribute(k_bluestein_bhat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_bluestein_bhat<<<1, 256, smem, st>>>(chirp, tw, bhat, T, planL);
        bt.chirp = chirp;
        bt.bhat = bhat;
        bt.L = L;
        plan = planL;
        threads = 256;
    }
    int rc = check_launch("hilbert setup", direct ? 1 : 3);
    if (rc) return rc;
    int per_sm = (int)((size_t)smem_max / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2048 / threads) per_sm = 2048 / threads;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > n_pairs) grid = n_pairs;
    if (direct) {
        LEIDA_CUDA_TRY(cudaFuncSetAttribute(k_hilbert_phase<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_hilbert_phase<false><<<(unsigned)grid, threads, smem, st>>>(x, n_series, T, x_pitch, phase, phase_pitch, t_begin,
                                                                      t_count, tw, plan, bt);
    } else {
        LEIDA_CUDA_TRY(cudaFuncSetAttribute(k_hilbert_phase<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_hilbert_phase<true><<<(unsigned)grid, threads, smem, st>>>(x, n_series, T, x_pitch, phase, phase_pitch, t_begin,
                                                                     t_count, tw, plan, bt);
    }
    return check_launch("k_hilbert_phase");
}

extern "C" int leida_phase_coherence_f64(const double* phase, int N, int T, int64_t phase_pitch, double* dfc,
                                         leida_stream_t stream) {
    LEIDA_REQUIRE(phase && dfc, "null pointer");
    LEIDA_REQUIRE(N >= 1 && T >= 3, "N >= 1 and T >= 3");
    LEIDA_REQUIRE(phase_pitch >= T, "phase_pitch >= T");
    const int Tp = T - 2;
    dim3 grid((Tp + 127) / 128, N);
    k_phase_coherence<<<grid, 128, 0, (cudaStream_t)stream>>>(phase, N, T, phase_pitch, dfc);
    return check_launch("k_phase_coherence");
}

extern "C" int leida_eigvec_rank2_f64(const double* phase, int64_t n_subjects, int N, int n_vol, int64_t subj_stride,
                                      int64_t row_pitch, double* lei, int64_t lei_pitch, float* xf, int64_t xf_pitch,
                                      leida_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_REQUIRE(phase, "null pointer");
    LEIDA_REQUIRE(lei || xf, "at least one output");
    LEIDA_REQUIRE(n_subjects >= 0 && N >= 1 && n_vol >= 0, "sizes");
    LEIDA_REQUIRE(n_subjects <= 65535, "n_subjects <= 65535 per call");
    LEIDA_REQUIRE(row_pitch >= n_vol, "row_pitch >= n_vol");
    LEIDA_REQUIRE(!lei || lei_pitch >= N, "lei_pitch >= N");
    LEIDA_REQUIRE(!xf || xf_pitch >= N, "xf_pitch >= N");
    if (n_subjects == 0 || n_vol == 0) return LEIDA_OK;
    // largest tile of volumes whose cos/sin fit comfortably (two CTAs per SM when possible)
    const size_t budget = 100 * 1024;
    const size_t per_t = (size_t)2 * (N | 1) * sizeof(double);
    if (16 * per_t + 9000 <= budget) return launch_eigvec<16>(phase, n_subjects, N, n_vol, subj_stride, row_pitch, lei, lei_pitch, xf, xf_pitch, st);
    if (8 * per_t + 9000 <= budget) return launch_eigvec<8>(phase, n_subjects, N, n_vol, subj_stride, row_pitch, lei, lei_pitch, xf, xf_pitch, st);
    if (4 * per_t + 9000 <= (size_t)max_smem_optin()) return launch_eigvec<4>(phase, n_subjects, N, n_vol, subj_stride, row_pitch, lei, lei_pitch, xf, xf_pitch, st);
    if (1 * per_t + 9000 <= (size_t)max_smem_optin()) return launch_eigvec<1>(phase, n_subjects, N, n_vol, subj_stride, row_pitch, lei, lei_pitch, xf, xf_pitch, st);
    set_error("N=%d too large for the rank-2 eigenvector kernel", N);
    return LEIDA_ERR_UNSUPPORTED;
}

// Fused stage 1+2: phases of a chunk of subjects go to the workspace (trimmed to the T-2 used
// volumes, pitch rounded up to 8 doubles so tiles are sector aligned) and are consumed while still
// L2-resident.
static inline int64_t trimmed_pitch(int T) { return ((int64_t)(T - 2) + 7) / 8 * 8; }
static inline int64_t fused_chunk_subjects(int64_t S, int N, int T) {
    const int64_t per_subject = (int64_t)N * trimmed_pitch(T) * 8;
    int64_t c = (48LL << 20) / (per_subject > 0 ? per_subject : 1);   // ~48 MB of phases per chunk (L2 = 126 MB)
    if (c < 1) c = 1;
    if (c > S) c = S;
    if (c > 65535) c = 65535;
    return c;
}

extern "C" size_t leida_leading_eigvec_workspace_bytes(int64_t n_subjects, int N, int T) {
    if (n_subjects < 1 || N < 1 || T < 3) return 0;
    const int64_t c = fused_chunk_subjects(n_subjects, N, T);
    size_t hil = (leida_hilbert_workspace_bytes(T) + 255) / 256 * 256;
    return hil + (size_t)c * N * trimmed_pitch(T) * 8;
}

extern "C" int leida_leading_eigvec_f64(const double* x, int64_t n_subjects, int N, int T, double* lei, int64_t lei_pitch,
                                        float* xf, int64_t xf_pitch, void* workspace, size_t workspace_bytes,
                                        leida_stream_t stream) {
    LEIDA_REQUIRE(x && workspace, "null pointer");
    LEIDA_REQUIRE(N >= 1 && T >= 3 && n_subjects >= 0, "N >= 1, T >= 3");
    if (workspace_bytes < leida_leading_eigvec_workspace_bytes(n_subjects, N, T)) {
        set_error("fused eigenvector workspace too small");
        return LEIDA_ERR_WORKSPACE;
    }
    if (n_subjects == 0) return LEIDA_OK;
    const int64_t chunk = fused_chunk_subjects(n_subjects, N, T);
    const size_t hil = (leida_hilbert_workspace_bytes(T) + 255) / 256 * 256;
    double* ph = (double*)((char*)workspace + hil);
    const int64_t tp = trimmed_pitch(T);
    const int n_vol = T - 2;
    for (int64_t s0 = 0; s0 < n_subjects; s0 += chunk) {
        const int64_t ns = (n_subjects - s0 < chunk) ? (n_subjects - s0) : chunk;
        int rc = leida_hilbert_phase_f64(x + s0 * (int64_t)N * T, ns * N, T, T, ph, tp, 1, n_vol, workspace, hil, stream);
        if (rc) return rc;
        rc = leida_eigvec_rank2_f64(ph, ns, N, n_vol, (int64_t)N * tp, tp,
                                    lei ? lei + s0 * n_vol * lei_pitch : nullptr, lei_pitch,
                                    xf ? xf + s0 * n_vol * xf_pitch : nullptr, xf_pitch, stream);
        if (rc) return rc;
    }
    return LEIDA_OK;
}
