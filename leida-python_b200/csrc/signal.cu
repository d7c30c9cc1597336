// Stage 1 + 2 of the LEiDA hot path on sm_100a:
//   * batched Hilbert phase (shared-memory mixed-radix / Bluestein FFT in fp64)
//       replaces pyleida/signal_tools/_signal_tools.py:24-44 (scipy.signal.hilbert + np.angle)
//   * phase coherence (standalone writer)          _signal_tools.py:137-183
//   * rank-2 closed-form leading eigenvector       _signal_tools.py:137-224 (dFC never materialised)
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace leida {

// ------------------------------------------------------------------------------------------------
// FFT machinery.  Stockham autosort, decimation in frequency, one CTA per transform, data in shared
// memory as double2.  Twiddles tw[k] = exp(-2 pi i k / n) live in global memory (L1-resident).
// ------------------------------------------------------------------------------------------------
// A pass is a Stockham step of radix r1*r2: r2 == 1 is a plain radix (2, 3, 4, 5 or an odd prime up to 31); r2 > 1 is
// a composite radix whose r1*r2-point butterfly is done in registers (Cooley-Tukey r1 x r2 with internal twiddles), so
// T = 1200 = (5*3) * 5 * (4*4) needs three trips through shared memory instead of five.
struct FftPlan {
    int n;
    int nfac;
    unsigned char r1[20];
    unsigned char r2[20];
    // internal twiddles of the composite passes: wint[slot(f)][m] = exp(-2 pi i m / (r1 r2)), m <= (r1-1)(r2-1) <= 9.
    // Kernel parameters live in the constant bank and the index is a compile-time constant, so these cost nothing.
    unsigned char slot[20];
    double2 wint[6][10];
};

static bool make_plan(int n, FftPlan& p) {
    p.n = n;
    p.nfac = 0;
    int small[32], ns = 0, big[32], nbig = 0;
    int m = n;
    while (m % 5 == 0) { small[ns++] = 5; m /= 5; }
    while (m % 4 == 0) { small[ns++] = 4; m /= 4; }
    while (m % 3 == 0) { small[ns++] = 3; m /= 3; }
    while (m % 2 == 0) { small[ns++] = 2; m /= 2; }
    for (int f = 7; f <= 31 && m > 1; f += 2)
        while (m % f == 0) { big[nbig++] = f; m /= f; }
    if (m != 1) return false;
    // pair the small factors (descending) greedily: the largest remaining factor takes the largest partner that keeps
    // the product <= 16 (the register budget of a composite butterfly)
    for (int i = 0; i < ns; ++i)
        for (int j = i + 1; j < ns; ++j)
            if (small[j] > small[i]) { int t = small[i]; small[i] = small[j]; small[j] = t; }
    bool used[32] = {false};
    for (int i = 0; i < ns; ++i) {
        if (used[i]) continue;
        used[i] = true;
        int partner = -1;
        for (int j = i + 1; j < ns; ++j)
            if (!used[j] && small[i] * small[j] <= 16) { partner = j; break; }
        p.r1[p.nfac] = (unsigned char)small[i];
        p.r2[p.nfac] = 1;
        if (partner >= 0) { used[partner] = true; p.r2[p.nfac] = (unsigned char)small[partner]; }
        ++p.nfac;
    }
    for (int i = 0; i < nbig; ++i) { p.r1[p.nfac] = (unsigned char)big[i]; p.r2[p.nfac] = 1; ++p.nfac; }
    // at most 6 distinct composite radices exist (16, 15, 12, 10, 9, 8, 6 minus what n admits); passes share slots
    int n_slots = 0, slot_r[6];
    for (int f = 0; f < p.nfac; ++f) {
        p.slot[f] = 0;
        if (p.r2[f] == 1) continue;
        const int R = p.r1[f] * p.r2[f];
        int sl = -1;
        for (int i = 0; i < n_slots; ++i) if (slot_r[i] == R) sl = i;
        if (sl < 0) {
            if (n_slots == 6) { p.r2[f] = 1; /* cannot happen: split back into a plain pass */ return false; }
            sl = n_slots++;
            slot_r[sl] = R;
            for (int m2 = 0; m2 < 10; ++m2) {
                const double a = -2.0 * 3.14159265358979323846 * (double)m2 / (double)R;
                p.wint[sl][m2] = make_double2(cos(a), sin(a));
            }
        }
        p.slot[f] = (unsigned char)sl;
    }
    return true;
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

// r1*r2-point DFT in registers, natural order in and out: X[b + R1 c] = sum_a w_R2^{ac} w_R^{ab} sum_d w_R1^{db} x[a + R2 d].
// w_R^m = wint[m] (FftPlan::wint, constant bank).
template <int R1, int R2>
__device__ __forceinline__ void butterfly2(double2* v, const double2* __restrict__ wint) {
    constexpr int R = R1 * R2;
    double2 y[R];
#pragma unroll
    for (int a = 0; a < R2; ++a) {
        double2 t[R1];
#pragma unroll
        for (int d = 0; d < R1; ++d) t[d] = v[a + R2 * d];
        butterfly<R1>(t);
#pragma unroll
        for (int b = 0; b < R1; ++b) y[a * R1 + b] = (a * b) ? cmul(t[b], wint[a * b]) : t[b];
    }
#pragma unroll
    for (int b = 0; b < R1; ++b) {
        double2 t[R2];
#pragma unroll
        for (int a = 0; a < R2; ++a) t[a] = y[a * R1 + b];
        butterfly<R2>(t);
#pragma unroll
        for (int c = 0; c < R2; ++c) v[b + R1 * c] = t[c];
    }
}

// One Stockham pass of radix R1*R2 over `batch` transforms held side by side in shared memory (transform i at
// in + i * stride): batch * n/R butterflies spread over the CTA.
// w^j for j = 1..R-1 from w = w^1 with multiplication depth <= 4 (w^2, w^4, w^8 by squaring, the rest as products of
// those): one table load per butterfly instead of R-1 scattered 16-byte gathers, at ~4 ulp.
template <int R>
__device__ __forceinline__ void twiddle_powers(double2 w, double2* pw) {
    pw[1] = w;
#pragma unroll
    for (int j = 2; j < R; ++j) {
        const int hb = j >= 8 ? 8 : j >= 4 ? 4 : 2;       // highest power of two <= j
        pw[j] = (j == hb) ? cmul(pw[j / 2], pw[j / 2]) : cmul(pw[hb], pw[j - hb]);
    }
}

template <int R1, int R2>
__device__ __forceinline__ void fft_pass(const double2* __restrict__ in, double2* __restrict__ out,
                                         const double2* __restrict__ tw, const double2* __restrict__ wint, int n, int len,
                                         int s, int batch, int stride) {
    constexpr int R = R1 * R2;
    const int m = len / R, nb = n / R, tstep = n / len;
    // idx / nb and i / s by float reciprocals (exact for operands < 2^20; the sizes here are < 2^14)
    const float inv_nb = 1.0f / (float)nb, inv_s = 1.0f / (float)s;
    for (int idx = threadIdx.x; idx < nb * batch; idx += blockDim.x) {
        const int bi = batch == 1 ? 0 : (int)(((float)idx + 0.5f) * inv_nb);
        const int i = idx - bi * nb;
        const int p = s == 1 ? i : (m == 1 ? 0 : (int)(((float)i + 0.5f) * inv_s));
        const int q = i - p * s;
        const double2* ib = in + (size_t)bi * stride;
        double2 v[R];
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = ib[q + s * (p + k * m)];
        if constexpr (R2 == 1) butterfly<R1>(v);
        else butterfly2<R1, R2>(v, wint);
        double2* o = out + (size_t)bi * stride + q + s * (R * p);
        o[0] = v[0];
        if (p == 0) {
#pragma unroll
            for (int j = 1; j < R; ++j) o[s * j] = v[j];
        } else {
            double2 pw[R];
            twiddle_powers<R>(__ldg(&tw[p * tstep]), pw);
#pragma unroll
            for (int j = 1; j < R; ++j) o[s * j] = cmul(v[j], pw[j]);
        }
    }
}

// Generic odd prime radix (7..31): O(r^2) butterfly from the twiddle table.
__device__ __noinline__ void fft_pass_generic(const double2* __restrict__ in, double2* __restrict__ out,
                                              const double2* __restrict__ tw, int n, int len, int s, int r, int batch,
                                              int stride) {
    const int m = len / r, nb = n / r, tstep = n / len, rstep = n / r;
    for (int idx = threadIdx.x; idx < nb * batch; idx += blockDim.x) {
        const int bi = idx / nb, i = idx - bi * nb;
        const int p = i / s, q = i - p * s;
        const double2* ib = in + (size_t)bi * stride;
        double2* ob = out + (size_t)bi * stride;
        double2 v[31];
        for (int k = 0; k < r; ++k) v[k] = ib[q + s * (p + k * m)];
        for (int j = 0; j < r; ++j) {
            double2 acc = v[0];
            for (int k = 1; k < r; ++k) acc = cadd(acc, cmul(v[k], __ldg(&tw[((j * k) % r) * rstep])));
            ob[q + s * (r * p + j)] = cmul(acc, __ldg(&tw[p * j * tstep]));
        }
    }
}

// Forward DFT of `batch` transforms of plan.n points (transform i at src + i * stride, scratch at dst + i * stride);
// on return `src` points at the results.
__device__ __forceinline__ void fft_forward(double2*& src, double2*& dst, const double2* __restrict__ tw,
                                            const FftPlan& plan, int batch = 1, int stride = 0) {
    int len = plan.n, s = 1;
    for (int f = 0; f < plan.nfac; ++f) {
        const int r1 = plan.r1[f], r2 = plan.r2[f];
        switch (r1 * 8 + r2) {
            case 4 * 8 + 4: fft_pass<4, 4>(src, dst, tw, plan.wint[plan.slot[f]], plan.n, len, s, batch, stride); break;
            case 5 * 8 + 3: fft_pass<5, 3>(src, dst, tw, plan.wint[plan.slot[f]], plan.n, len, s, batch, stride); break;
            case 5 * 8 + 2: fft_pass<5, 2>(src, dst, tw, plan.wint[plan.slot[f]], plan.n, len, s, batch, stride); break;
            case 4 * 8 + 3: fft_pass<4, 3>(src, dst, tw, plan.wint[plan.slot[f]], plan.n, len, s, batch, stride); break;
            case 4 * 8 + 2: fft_pass<4, 2>(src, dst, tw, plan.wint[plan.slot[f]], plan.n, len, s, batch, stride); break;
            case 3 * 8 + 3: fft_pass<3, 3>(src, dst, tw, plan.wint[plan.slot[f]], plan.n, len, s, batch, stride); break;
            case 3 * 8 + 2: fft_pass<3, 2>(src, dst, tw, plan.wint[plan.slot[f]], plan.n, len, s, batch, stride); break;
            case 5 * 8 + 1: fft_pass<5, 1>(src, dst, tw, plan.wint[plan.slot[f]], plan.n, len, s, batch, stride); break;
            case 4 * 8 + 1: fft_pass<4, 1>(src, dst, tw, plan.wint[plan.slot[f]], plan.n, len, s, batch, stride); break;
            case 3 * 8 + 1: fft_pass<3, 1>(src, dst, tw, plan.wint[plan.slot[f]], plan.n, len, s, batch, stride); break;
            case 2 * 8 + 1: fft_pass<2, 1>(src, dst, tw, plan.wint[plan.slot[f]], plan.n, len, s, batch, stride); break;
            default: fft_pass_generic(src, dst, tw, plan.n, len, s, r1, batch, stride); break;
        }
        __syncthreads();
        double2* t = src; src = dst; dst = t;
        len /= r1 * r2;
        s *= r1 * r2;
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

// (cos, sin) of atan2(h, x) without the angle: (x, h) / hypot(x, h); atan2(0, 0) = 0 -> (1, 0).
__device__ __forceinline__ double2 unit2(double x, double h) {
    const double r2 = x * x + h * h;
    if (r2 > 1e-280 && r2 < 1e280) {
        const double inv = rsqrt(r2);                   // <= 1 ulp (CUDA math API): cheaper than sqrt + divide
        return make_double2(x * inv, h * inv);
    }
    const double r = hypot(x, h);                       // zero, subnormal or huge: the slow, safe way
    if (!(r > 0.0)) return make_double2(1.0, 0.0);
    return make_double2(x / r, h / r);
}

// Hilbert transform of two real series per complex FFT: z = x1 + i x2; A(z) = ifft(h fft(z)) =
// (x1 - H x2) + i (x2 + H x1)  =>  H x1 = Im A - x2,  H x2 = x1 - Re A.  `batch` such pairs per CTA pass through the
// FFT side by side (more butterflies per barrier).
//   CS == false: out = phase = atan2(H x, x)                          (hilbert_phase, _signal_tools.py:39-43)
//   CS == true : out = (cos, sin) of that phase = (x, H x) / |(x, H x)| as double2 -- what the rank-2 eigenvector
//                needs; no angle is ever formed on the fused path (no atan2 here, no sincos in k_eigvec_rank2).
template <bool BLUESTEIN, bool CS>
__global__ void __launch_bounds__(256, 2)
k_hilbert(const double* __restrict__ x, long long n_series, int T, long long x_pitch,
                          double* __restrict__ out, long long out_pitch, int t_begin, int t_count,
                          const double2* __restrict__ tw, FftPlan plan, BluesteinTables bt, int batch) {
    extern __shared__ double2 sm[];
    const int bufn = BLUESTEIN ? bt.L : T;
    const int stride = 2 * bufn;                       // per transform: two buffers
    const long long n_pairs = (n_series + 1) / 2;
    const long long n_groups = (n_pairs + batch - 1) / batch;
    const int half = (T + 1) / 2;
    const double invT = 1.0 / (double)T;
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long long pair0 = g * batch;
        const int nb = (int)((n_pairs - pair0) < batch ? (n_pairs - pair0) : batch);
        double2* src = sm;
        double2* dst = sm + bufn;
        for (int idx = threadIdx.x; idx < batch * T; idx += blockDim.x) {
            const int bi = batch == 1 ? 0 : idx / T, t = idx - bi * T;
            double2 v = make_double2(0.0, 0.0);
            if (bi < nb) {
                const long long s1 = 2 * (pair0 + bi);
                const double* x1 = x + s1 * x_pitch;
                v.x = x1[t];
                if (s1 + 1 < n_series) v.y = x1[x_pitch + t];
            }
            src[(size_t)bi * stride + t] = v;
        }
        __syncthreads();
        if (BLUESTEIN) dft_bluestein(src, dst, T, tw, plan, bt);
        else fft_forward(src, dst, tw, plan, batch, stride);
        for (int idx = threadIdx.x; idx < batch * T; idx += blockDim.x) {
            const int bi = batch == 1 ? 0 : idx / T, k = idx - bi * T;
            const double h = (k == 0 || (2 * k == T)) ? 1.0 : (k < half ? 2.0 : 0.0);
            double2* e = src + (size_t)bi * stride + k;
            const double2 v = *e;
            *e = make_double2(h * v.x, -h * v.y);
        }
        __syncthreads();
        if (BLUESTEIN) dft_bluestein(src, dst, T, tw, plan, bt);
        else fft_forward(src, dst, tw, plan, batch, stride);
        for (int idx = threadIdx.x; idx < nb * t_count; idx += blockDim.x) {
            const int bi = batch == 1 ? 0 : idx / t_count, i = idx - bi * t_count;
            const int t = t_begin + i;
            const long long s1 = 2 * (pair0 + bi);
            const bool has2 = s1 + 1 < n_series;
            const double* x1 = x + s1 * x_pitch;
            const double2 r = src[(size_t)bi * stride + t];
            const double are = r.x * invT, aim = -r.y * invT;
            const double v1 = x1[t], v2 = has2 ? x1[x_pitch + t] : 0.0;
            const double h1 = aim - v2, h2 = v1 - are;
            if (CS) {
                double2* o = reinterpret_cast<double2*>(out) + s1 * out_pitch + i;
                o[0] = unit2(v1, h1);
                if (has2) o[out_pitch] = unit2(v2, h2);
            } else {
                double* o = out + s1 * out_pitch + i;
                o[0] = atan2(h1, v1);
                if (has2) o[out_pitch] = atan2(h2, v2);
            }
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
// CS == true: `phase` holds (cos, sin) pairs (double2; strides in double2 units) written by k_hilbert<.., true>.
template <int TT, bool CS>
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
    const double* ph = phase + s * subj_stride * (CS ? 2 : 1);
    const bool live = (v0 + tt) < n_vol;

    double c2 = 0.0, s2 = 0.0;
    for (int e = tid; e < N * TT; e += 256) {
        const int j = e / TT;
        double sn = 0.0, cs = 0.0;
        if (live) {
            if (CS) {
                const double2 u = reinterpret_cast<const double2*>(ph)[(long long)j * row_pitch + v0 + tt];
                cs = u.x;
                sn = u.y;
            } else {
                sincos(ph[(long long)j * row_pitch + v0 + tt], &sn, &cs);
            }
        }
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

template <int TT, bool CS>
static int launch_eigvec(const double* phase, int64_t S, int N, int n_vol, int64_t subj_stride, int64_t row_pitch,
                         double* lei, int64_t lei_pitch, float* xf, int64_t xf_pitch, cudaStream_t st) {
    const size_t smem = ((size_t)2 * TT * (N | 1) + 4 * 256 + 4 * TT) * sizeof(double);
    LEIDA_CUDA_TRY(cudaFuncSetAttribute(k_eigvec_rank2<TT, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((n_vol + TT - 1) / TT, (unsigned)S);
    k_eigvec_rank2<TT, CS><<<grid, 256, smem, st>>>(phase, N, n_vol, subj_stride, row_pitch, lei, lei_pitch, xf, xf_pitch);
    return check_launch("k_eigvec_rank2");
}

// phase (CS == false) or (cos, sin) pairs (CS == true) -> eigenvector rows; picks the volume tile by shared memory.
template <bool CS>
static int eigvec_rank2(const double* phase, int64_t n_subjects, int N, int n_vol, int64_t subj_stride, int64_t row_pitch,
                        double* lei, int64_t lei_pitch, float* xf, int64_t xf_pitch, cudaStream_t st) {
    // largest tile of volumes whose cos/sin fit comfortably (two CTAs per SM when possible)
    const size_t budget = 100 * 1024;
    const size_t per_t = (size_t)2 * (N | 1) * sizeof(double);
    if (16 * per_t + 9000 <= budget) return launch_eigvec<16, CS>(phase, n_subjects, N, n_vol, subj_stride, row_pitch, lei, lei_pitch, xf, xf_pitch, st);
    if (8 * per_t + 9000 <= budget) return launch_eigvec<8, CS>(phase, n_subjects, N, n_vol, subj_stride, row_pitch, lei, lei_pitch, xf, xf_pitch, st);
    if (4 * per_t + 9000 <= (size_t)max_smem_optin()) return launch_eigvec<4, CS>(phase, n_subjects, N, n_vol, subj_stride, row_pitch, lei, lei_pitch, xf, xf_pitch, st);
    if (1 * per_t + 9000 <= (size_t)max_smem_optin()) return launch_eigvec<1, CS>(phase, n_subjects, N, n_vol, subj_stride, row_pitch, lei, lei_pitch, xf, xf_pitch, st);
    set_error("N=%d too large for the rank-2 eigenvector kernel", N);
    return LEIDA_ERR_UNSUPPORTED;
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

// Threads per CTA for `batch` transforms per CTA: the count (multiple of 32, 64..256) that wastes the fewest thread
// slots over the plan's passes (a pass has batch * n / radix butterflies).
static int pick_threads(const FftPlan& plan, int batch) {
    int best = 256;
    double best_u = -1.0;
    for (int th = 64; th <= 256; th += 32) {
        double work = 0.0, slots = 0.0;
        for (int f = 0; f < plan.nfac; ++f) {
            const int tasks = batch * (plan.n / (plan.r1[f] * plan.r2[f]));
            work += tasks;
            slots += (double)((tasks + th - 1) / th) * th;
        }
        const double u = work / slots + 1e-4 * th;         // ties: more threads (more loads in flight)
        if (u > best_u) { best_u = u; best = th; }
    }
    return best;
}

// cs == 0: out = phases (double, pitch in doubles); cs != 0: out = (cos, sin) pairs (double2, pitch in double2 units).
static int hilbert_launch(const double* x, int64_t n_series, int T, int64_t x_pitch, double* out, int64_t out_pitch,
                          int t_begin, int t_count, void* workspace, size_t workspace_bytes, int cs, cudaStream_t st) {
    LEIDA_REQUIRE(x && out && workspace, "null pointer");
    LEIDA_REQUIRE(n_series >= 0 && T >= 1, "n_series >= 0 and T >= 1");
    LEIDA_REQUIRE(x_pitch >= T, "x_pitch >= T");
    LEIDA_REQUIRE(t_begin >= 0 && t_count >= 0 && t_begin + t_count <= T, "volume window inside [0,T)");
    LEIDA_REQUIRE(out_pitch >= t_count, "phase_pitch >= t_count");
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
    int threads, batch = 1;
    if (direct) {
        const size_t per = (size_t)2 * T * sizeof(double2);
        if ((long long)per > smem_max) {
            set_error("T=%d too long for the shared-memory FFT (max %d)", T, smem_max / 32);
            return LEIDA_ERR_UNSUPPORTED;
        }
        // One transform pair per CTA and several small CTAs per SM (measured, profiles/r02_stage12.md): the kernel is
        // bound by latency (input loads, barriers between passes), which other resident CTAs hide better than a wider
        // CTA does; short transforms are batched so that a CTA still has a warp's worth of butterflies per pass.
        while (batch < 8 && (batch + 1) * per <= (size_t)16 << 10 && (long long)(batch + 1) * sm_count() * 4 <= n_pairs) ++batch;
        if (const char* e = getenv("LEIDA_HILBERT_BATCH")) {          // tuning experiments only
            const int b = atoi(e);
            if (b >= 1 && (size_t)b * per <= (size_t)smem_max) batch = b;
        }
        smem = batch * per;
        k_twiddles<<<(T + 255) / 256, 256, 0, st>>>(tw, T);
        threads = pick_threads(plan, batch);
        if (batch == 1 && T >= 1024 && threads < 128) threads = 128;
        if (const char* e = getenv("LEIDA_HILBERT_THREADS")) {
            const int t = atoi(e);
            if (t >= 32 && t <= 256 && t % 32 == 0) threads = t;
        }
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
    // Persistent CTAs, exactly as many as are resident at once: shared memory would admit five 128-thread CTAs per SM at
    // T = 1200, the register file four -- a grid sized by shared memory alone ran its last fifth as a second wave
    // (4 group times instead of 3 for the 1600 pairs of a config-3 chunk).
    const long long n_groups = (n_pairs + batch - 1) / batch;
#define LEIDA_HILBERT_LAUNCH(BL, CSV)                                                                                     \
    do {                                                                                                                  \
        LEIDA_CUDA_TRY(cudaFuncSetAttribute(k_hilbert<BL, CSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        int per_sm = 0;                                                                                                   \
        LEIDA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_hilbert<BL, CSV>, threads, smem));        \
        if (per_sm < 1) per_sm = 1;                                                                                       \
        long long grid = (long long)sm_count() * per_sm;                                                                  \
        if (grid > n_groups) grid = n_groups;                                                                             \
        k_hilbert<BL, CSV><<<(unsigned)grid, threads, smem, st>>>(x, n_series, T, x_pitch, out, out_pitch, t_begin,     \
                                                                  t_count, tw, plan, bt, batch);                         \
    } while (0)
    if (direct) { if (cs) LEIDA_HILBERT_LAUNCH(false, true); else LEIDA_HILBERT_LAUNCH(false, false); }
    else { if (cs) LEIDA_HILBERT_LAUNCH(true, true); else LEIDA_HILBERT_LAUNCH(true, false); }
#undef LEIDA_HILBERT_LAUNCH
    return check_launch("k_hilbert");
}

extern "C" int leida_hilbert_phase_f64(const double* x, int64_t n_series, int T, int64_t x_pitch, double* phase,
                                       int64_t phase_pitch, int t_begin, int t_count, void* workspace,
                                       size_t workspace_bytes, leida_stream_t stream) {
    return hilbert_launch(x, n_series, T, x_pitch, phase, phase_pitch, t_begin, t_count, workspace, workspace_bytes, 0,
                          (cudaStream_t)stream);
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
    return eigvec_rank2<false>(phase, n_subjects, N, n_vol, subj_stride, row_pitch, lei, lei_pitch, xf, xf_pitch, st);
}

// Fused stage 1+2: the (cos, sin) pairs of a chunk of subjects go to the workspace (trimmed to the T-2 used volumes,
// pitch rounded up to 8 so tiles are sector aligned) and are consumed by the eigenvector kernel while still L2-resident.
static inline int64_t trimmed_pitch(int T) { return ((int64_t)(T - 2) + 7) / 8 * 8; }
static inline int64_t fused_chunk_subjects(int64_t S, int N, int T) {
    const int64_t per_subject = (int64_t)N * trimmed_pitch(T) * 16;
    int64_t c = (64LL << 20) / (per_subject > 0 ? per_subject : 1);   // ~64 MB of pairs per chunk (L2 = 126 MB)
    if (c < 1) c = 1;
    if (c > S) c = S;
    if (c > 65535) c = 65535;
    return c;
}

extern "C" size_t leida_leading_eigvec_workspace_bytes(int64_t n_subjects, int N, int T) {
    if (n_subjects < 1 || N < 1 || T < 3) return 0;
    const int64_t c = fused_chunk_subjects(n_subjects, N, T);
    size_t hil = (leida_hilbert_workspace_bytes(T) + 255) / 256 * 256;
    return hil + (size_t)c * N * trimmed_pitch(T) * 16;
}

extern "C" int leida_leading_eigvec_f64(const double* x, int64_t n_subjects, int N, int T, double* lei, int64_t lei_pitch,
                                        float* xf, int64_t xf_pitch, void* workspace, size_t workspace_bytes,
                                        leida_stream_t stream) {
    LEIDA_REQUIRE(x && workspace, "null pointer");
    LEIDA_REQUIRE(N >= 1 && T >= 3 && n_subjects >= 0, "N >= 1, T >= 3");
    LEIDA_REQUIRE(lei || xf, "at least one output");
    LEIDA_REQUIRE(!lei || lei_pitch >= N, "lei_pitch >= N");
    LEIDA_REQUIRE(!xf || xf_pitch >= N, "xf_pitch >= N");
    if (workspace_bytes < leida_leading_eigvec_workspace_bytes(n_subjects, N, T)) {
        set_error("fused eigenvector workspace too small");
        return LEIDA_ERR_WORKSPACE;
    }
    if (n_subjects == 0) return LEIDA_OK;
    const int64_t chunk = fused_chunk_subjects(n_subjects, N, T);
    const size_t hil = (leida_hilbert_workspace_bytes(T) + 255) / 256 * 256;
    double* cs = (double*)((char*)workspace + hil);
    const int64_t tp = trimmed_pitch(T);
    const int n_vol = T - 2;
    for (int64_t s0 = 0; s0 < n_subjects; s0 += chunk) {
        const int64_t ns = (n_subjects - s0 < chunk) ? (n_subjects - s0) : chunk;
        int rc = hilbert_launch(x + s0 * (int64_t)N * T, ns * N, T, T, cs, tp, 1, n_vol, workspace, hil, 1, (cudaStream_t)stream);
        if (rc) return rc;
        rc = eigvec_rank2<true>(cs, ns, N, n_vol, (int64_t)N * tp, tp, lei ? lei + s0 * n_vol * lei_pitch : nullptr, lei_pitch,
                                xf ? xf + s0 * n_vol * xf_pitch : nullptr, xf_pitch, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return LEIDA_OK;
}
