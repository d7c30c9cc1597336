// Stage 0 (optional, in front of the hot path): clean_signals -- pyleida/signal_tools/_signal_tools.py:66-133, which
// wraps nilearn.signal.clean (nilearn 0.9.1: signal.py clean -> _detrend(type='linear') -> butterworth ->
// _standardize).  nilearn is a third-party dependency that is neither installed nor vendored, so this follows its
// published algorithm for the options clean_signals exposes (detrend, standardize in {zscore, psc, False},
// filter_type in {butterworth, None}); the Butterworth design itself (scipy.signal.butter -> b, a, and lfilter_zi)
// is scalar host work and comes in as arguments.  SURVEY 8f-4; parity is unpinned against nilearn, pinned against the
// restatement in the test oracle (which runs the real scipy.signal.filtfilt).
//
// The IIR recursion is sequential in time, so the parallel axis is the series: one thread per series, all arrays in a
// time-major scratch layout W[t][series] so that every load and store of a warp is one coalesced line.
//   k_clean_transpose (in)  -> k_clean_series (mean / linear detrend, filtfilt = forward + backward pass of scipy's
//   direct-form-II-transposed lfilter over the odd-extended series with lfilter_zi initial states, standardize)
//   -> k_clean_transpose (out).
#include "common.cuh"
#include <math.h>

namespace leida {

constexpr int CLEAN_MAX_COEF = 12;        // order-5 band-pass: 11 coefficients

struct CleanFilter {
    int nc;                                // number of coefficients, 0 = no filter
    double b[CLEAN_MAX_COEF];
    double a[CLEAN_MAX_COEF];
    double zi[CLEAN_MAX_COEF];
};

// 32 x 32 tile transpose through shared memory: src (rows, cols) pitch sp -> dst (cols, rows) pitch dp.
__global__ void k_clean_transpose(const double* __restrict__ src, long long rows, long long cols, long long sp,
                                  double* __restrict__ dst, long long dp) {
    __shared__ double tile[32][33];
    const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = src[r * sp + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[c * dp + r] = tile[threadIdx.x][i];
    }
}

// One step of scipy.signal.lfilter's direct form II transposed (a[0] = 1), in its order of operations and without
// fused multiply-adds: y = z0 + b0 x;  z_i = (z_{i+1} + x b_{i+1}) - y a_{i+1};  z_last = x b_last - y a_last.
__device__ __forceinline__ double lfilter_step(const CleanFilter& f, double* z, double x) {
    const double y = __dadd_rn(z[0], __dmul_rn(f.b[0], x));
#pragma unroll
    for (int i = 0; i < CLEAN_MAX_COEF - 2; ++i)
        if (i < f.nc - 2) z[i] = __dsub_rn(__dadd_rn(z[i + 1], __dmul_rn(x, f.b[i + 1])), __dmul_rn(y, f.a[i + 1]));
#pragma unroll
    for (int i = 0; i < CLEAN_MAX_COEF - 1; ++i)
        if (i == f.nc - 2) z[i] = __dsub_rn(__dmul_rn(x, f.b[i + 1]), __dmul_rn(y, f.a[i + 1]));
    return y;
}

// W0: (T, ns) time-major input, overwritten with the result.  W1: (T + 2 padlen, ns) scratch for the forward pass.
__global__ void __launch_bounds__(128)
k_clean_series(double* __restrict__ W0, double* __restrict__ W1, long long ns, long long pitch, int T, int detrend,
               int standardize, double reg_mean, double reg_std, const __grid_constant__ CleanFilter f, int padlen) {
    const long long sidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= ns) return;
    double* x = W0 + sidx;
    const double eps = 2.220446049250313e-16;
    // mean of the raw series (nilearn keeps it for percent signal change)
    double sum = 0.0;
    for (int t = 0; t < T; ++t) sum += x[(long long)t * pitch];
    const double mean0 = sum / (double)T;
    if (detrend) {                                                    // _detrend(type='linear')
        double dot = 0.0;
        const bool use_std = !(reg_std < eps);
        for (int t = 0; t < T; ++t) {
            const double v = x[(long long)t * pitch] - mean0;
            double r = (double)t - reg_mean;
            if (use_std) r = r / reg_std;
            dot += r * v;
        }
        for (int t = 0; t < T; ++t) {
            double r = (double)t - reg_mean;
            if (use_std) r = r / reg_std;
            x[(long long)t * pitch] = (x[(long long)t * pitch] - mean0) - dot * r;
        }
    }
    if (f.nc > 0) {                                                   // scipy.signal.filtfilt(b, a, x): padtype='odd'
        double* y1 = W1 + sidx;
        const int L = T + 2 * padlen;
        const double x_first = x[0], x_last = x[(long long)(T - 1) * pitch];
        auto ext = [&](int e) -> double {
            if (e < padlen) return 2.0 * x_first - x[(long long)(padlen - e) * pitch];
            if (e < padlen + T) return x[(long long)(e - padlen) * pitch];
            return 2.0 * x_last - x[(long long)(T - 2 - (e - padlen - T)) * pitch];
        };
        double z[CLEAN_MAX_COEF];
        const double e0 = ext(0);
#pragma unroll
        for (int i = 0; i < CLEAN_MAX_COEF; ++i) z[i] = i < f.nc - 1 ? f.zi[i] * e0 : 0.0;
        for (int e = 0; e < L; ++e) y1[(long long)e * pitch] = lfilter_step(f, z, ext(e));
        const double yl = y1[(long long)(L - 1) * pitch];
#pragma unroll
        for (int i = 0; i < CLEAN_MAX_COEF; ++i) z[i] = i < f.nc - 1 ? f.zi[i] * yl : 0.0;
        for (int e = L - 1; e >= 0; --e) {
            const double y = lfilter_step(f, z, y1[(long long)e * pitch]);
            if (e >= padlen && e < padlen + T) x[(long long)(e - padlen) * pitch] = y;
        }
    }
    if (standardize == 1 && T > 1) {                                  // _standardize(standardize='zscore', detrend=False)
        double s1 = 0.0;
        for (int t = 0; t < T; ++t) s1 += x[(long long)t * pitch];
        const double m = s1 / (double)T;
        double s2 = 0.0;
        for (int t = 0; t < T; ++t) s2 += x[(long long)t * pitch] - m;
        const double m2 = s2 / (double)T;                             // numpy's std() takes the mean again
        double ss = 0.0;
        for (int t = 0; t < T; ++t) {
            const double d = (x[(long long)t * pitch] - m) - m2;
            ss += d * d;
        }
        double sd = sqrt(ss / (double)T);
        if (sd < eps) sd = 1.0;
        for (int t = 0; t < T; ++t) x[(long long)t * pitch] = (x[(long long)t * pitch] - m) / sd;
    } else if (standardize == 2 && T > 1) {                           // percent signal change
        const double add = detrend ? mean0 : 0.0;
        double s1 = 0.0;
        for (int t = 0; t < T; ++t) s1 += x[(long long)t * pitch] + add;
        const double m = s1 / (double)T;
        const bool invalid = fabs(m) < eps;
        for (int t = 0; t < T; ++t) {
            const double v = ((x[(long long)t * pitch] + add) - m) / fabs(m) * 100.0;
            x[(long long)t * pitch] = invalid ? 0.0 : v;
        }
    }
}

}  // namespace leida

using namespace leida;

static inline int64_t clean_chunk(int64_t n_series) { return n_series < (1 << 16) ? n_series : (1 << 16); }
static inline int64_t clean_pitch(int64_t ns) { return (ns + 31) / 32 * 32; }

extern "C" size_t leida_clean_workspace_bytes(int64_t n_series, int T, int n_coef) {
    if (n_series < 1 || T < 1) return 0;
    const int64_t p = clean_pitch(clean_chunk(n_series));
    const int padlen = n_coef > 0 ? 3 * n_coef : 0;
    return (size_t)p * ((size_t)T + (n_coef > 0 ? (size_t)T + 2 * padlen : 0)) * sizeof(double);
}

extern "C" int leida_clean_signals_f64(const double* x, int64_t n_series, int T, int64_t x_pitch, double* out,
                                       int64_t out_pitch, int detrend, int standardize, const double* b, const double* a,
                                       const double* zi, int n_coef, void* workspace, size_t workspace_bytes,
                                       leida_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_REQUIRE(x && out && workspace, "null pointer");
    LEIDA_REQUIRE(n_series >= 0 && T >= 1 && x_pitch >= T && out_pitch >= T, "sizes");
    LEIDA_REQUIRE(standardize >= 0 && standardize <= 2, "standardize in {0: none, 1: zscore, 2: psc}");
    LEIDA_REQUIRE(n_coef == 0 || (n_coef >= 2 && n_coef <= CLEAN_MAX_COEF && b && a && zi), "2 <= n_coef <= 12 with b, a, zi");
    const int padlen = n_coef > 0 ? 3 * n_coef : 0;
    LEIDA_REQUIRE(n_coef == 0 || T > padlen, "the series must be longer than filtfilt's padding (3 * n_coef)");
    if (workspace_bytes < leida_clean_workspace_bytes(n_series, T, n_coef)) {
        set_error("clean workspace too small");
        return LEIDA_ERR_WORKSPACE;
    }
    if (n_series == 0) return LEIDA_OK;
    CleanFilter f;
    f.nc = n_coef;
    for (int i = 0; i < CLEAN_MAX_COEF; ++i) {
        f.b[i] = i < n_coef ? b[i] / a[0] : 0.0;                      // scipy normalises by a[0]
        f.a[i] = i < n_coef ? a[i] / a[0] : 0.0;
        f.zi[i] = i < n_coef - 1 ? zi[i] : 0.0;
    }
    // numpy: regressor = arange(T) - mean; std = sqrt(sum(regressor ** 2)); (t - mean) is exact in fp64
    const double reg_mean = 0.5 * (double)(T - 1);
    double ss = 0.0;
    for (int t = 0; t < T; ++t) ss += ((double)t - reg_mean) * ((double)t - reg_mean);
    const double reg_std = sqrt(ss);
    const int64_t chunk = clean_chunk(n_series);
    const int64_t pitch = clean_pitch(chunk);
    double* W0 = (double*)workspace;
    double* W1 = W0 + (size_t)pitch * T;
    const dim3 tb(32, 8);
    for (int64_t s0 = 0; s0 < n_series; s0 += chunk) {
        const int64_t ns = n_series - s0 < chunk ? n_series - s0 : chunk;
        dim3 g_in((unsigned)((T + 31) / 32), (unsigned)((ns + 31) / 32));
        k_clean_transpose<<<g_in, tb, 0, st>>>(x + s0 * x_pitch, ns, T, x_pitch, W0, pitch);
        k_clean_series<<<(unsigned)((ns + 127) / 128), 128, 0, st>>>(W0, W1, ns, pitch, T, detrend, standardize, reg_mean,
                                                                     reg_std, f, padlen);
        dim3 g_out((unsigned)((ns + 31) / 32), (unsigned)((T + 31) / 32));
        k_clean_transpose<<<g_out, tb, 0, st>>>(W0, T, ns, pitch, out + s0 * out_pitch, out_pitch);
        const int rc = check_launch("clean_signals", 3);
        if (rc) return rc;
    }
    return LEIDA_OK;
}
