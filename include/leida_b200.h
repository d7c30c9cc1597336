/*
 * leida_b200.h -- C-ABI of the B200-native LEiDA hot path (libleida_b200.so).
 *
 * The reference (pyleida 1.0) is pure Python and has no FFI of its own; its "operator interface"
 * for this path is the Python call surface listed in SURVEY.md section 8b.  Each entry point below
 * names the reference function (file:line under /root/reference/pyleida) whose arithmetic it
 * replaces.  INTEGRATION.md shows the ctypes binding a pyleida maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host";
 *   - the caller owns all memory (inputs, outputs, workspaces); nothing here allocates or frees;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and
 *     keeps no global state besides a thread-local error string;
 *   - return value: 0 = ok, <0 = error (see LEIDA_ERR_*); leida_last_error() explains.
 *   - "pitch" arguments are in ELEMENTS of the array's type.
 */
#ifndef LEIDA_B200_H
#define LEIDA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LEIDA_OK                 0
#define LEIDA_ERR_INVALID       -1   /* bad argument */
#define LEIDA_ERR_UNSUPPORTED   -2   /* size outside what the kernels handle */
#define LEIDA_ERR_CUDA          -3   /* launch / runtime failure */
#define LEIDA_ERR_WORKSPACE     -4   /* workspace too small */

typedef void* leida_stream_t;

int         leida_version(void);
const char* leida_last_error(void);
/* number of SMs of the current device (grid sizing is derived from it) */
int         leida_device_sm_count(void);
/* kernels launched by this library in this process so far (monotonic counter, for bench.py) */
int64_t     leida_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Stage 0 (optional) -- clean_signals   (signal_tools/_signal_tools.py:66-133 -> nilearn.signal.clean, nilearn 0.9.1:
 * linear detrend, Butterworth band-pass applied forward and backward (scipy.signal.filtfilt), standardisation)
 * x: n_series rows of length T (row i at x + i*x_pitch) -> out, same layout (out == x is allowed).
 * detrend != 0: subtract the mean and the least-squares line.  standardize: 0 none, 1 z-score, 2 percent signal change.
 * b, a, zi: HOST arrays -- the n_coef transfer-function coefficients of scipy.signal.butter(5, Wn, btype, output='ba')
 * and the n_coef-1 steady-state initial conditions scipy.signal.lfilter_zi(b, a); the filter design is scalar host work
 * done by the caller.  n_coef = 0: no filtering; otherwise 2 <= n_coef <= 12 and T > 3*n_coef (filtfilt's odd padding).
 * One thread per series runs scipy's direct-form-II-transposed recursion over the odd-extended series in both
 * directions; data is held time-major in the workspace so that every access is coalesced.
 * ------------------------------------------------------------------------------------------ */
size_t leida_clean_workspace_bytes(int64_t n_series, int T, int n_coef);
int    leida_clean_signals_f64(const double* x, int64_t n_series, int T, int64_t x_pitch,
                               double* out, int64_t out_pitch, int detrend, int standardize,
                               const double* b, const double* a, const double* zi, int n_coef,
                               void* workspace, size_t workspace_bytes, leida_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Stage 1 -- hilbert_phase          (signal_tools/_signal_tools.py:24-44)
 * n_series real series of length T (row i at x + i*x_pitch) -> instantaneous phase in (-pi, pi].
 * Only volumes [t_begin, t_begin+t_count) are written, to phase + i*phase_pitch + (t - t_begin);
 * t_begin=0,t_count=T gives the reference's full (N,T) output.
 * Shared-memory Stockham FFT in fp64, two real series per complex transform; composite radices up to 16 are done in
 * registers (T = 1200 = 15 x 5 x 16: three passes), generic primes <= 31; other lengths use Bluestein.
 * Supported: T <= 7168 when T factors into primes <= 31, else T <= 2048.
 * ------------------------------------------------------------------------------------------ */
size_t leida_hilbert_workspace_bytes(int T);
int    leida_hilbert_phase_f64(const double* x, int64_t n_series, int T, int64_t x_pitch,
                               double* phase, int64_t phase_pitch, int t_begin, int t_count,
                               void* workspace, size_t workspace_bytes, leida_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Stage 2a -- phase_coherence       (signal_tools/_signal_tools.py:137-183)
 * phase: (N, T) with row pitch; dfc: (N, N, T-2) contiguous, t fastest; dFC[i,j,t] =
 * cos(theta_i(t+1) - theta_j(t+1)).
 * ------------------------------------------------------------------------------------------ */
int leida_phase_coherence_f64(const double* phase, int N, int T, int64_t phase_pitch,
                              double* dfc, leida_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Stage 2b -- get_eigenvectors(phase_coherence(.), n=1) without materialising dFC
 *                                   (signal_tools/_signal_tools.py:137-224)
 * phase points at the FIRST volume to use of subject 0 / ROI 0; subject s, ROI j, volume v lives
 * at phase + s*subj_stride + j*row_pitch + v, v in [0, n_vol).  Output row s*n_vol + v:
 *   lei (fp64, pitch lei_pitch, may be NULL) and xf (fp32, pitch xf_pitch >= N, may be NULL; the
 *   columns N..xf_pitch-1 are zero-filled -- the layout the k-means kernels want).
 * Unit 2-norm rows, reference sign rule (:215-220).
 * ------------------------------------------------------------------------------------------ */
int leida_eigvec_rank2_f64(const double* phase, int64_t n_subjects, int N, int n_vol,
                           int64_t subj_stride, int64_t row_pitch,
                           double* lei, int64_t lei_pitch, float* xf, int64_t xf_pitch,
                           leida_stream_t stream);

/* Fused stage 1+2 for the Leida pipeline (_leida.py:287-289): x (S, N, T) contiguous fp64 ->
 * eigenvector rows.  No angle is formed on this path: the Hilbert kernel writes (cos, sin) = (x, Hx)/|(x, Hx)| of a
 * chunk of subjects to `workspace` (L2-sized chunks) and the rank-2 kernel consumes them (no atan2, no sincos). */
size_t leida_leading_eigvec_workspace_bytes(int64_t n_subjects, int N, int T);
int    leida_leading_eigvec_f64(const double* x, int64_t n_subjects, int N, int T,
                                double* lei, int64_t lei_pitch, float* xf, int64_t xf_pitch,
                                void* workspace, size_t workspace_bytes, leida_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Stage 2b (general path) -- get_eigenvectors(dFC, n=1) for an arbitrary symmetric dFC
 *                                   (signal_tools/_signal_tools.py:185-224, eigs(...,'LM') at :213)
 * dfc: (N, N, n_vol) contiguous, t fastest.  Blocked subspace iteration + Rayleigh-Ritz, all
 * volumes advanced together.  out: (n_vol, N) fp64, unit rows, sign rule applied.
 * info (device, 4 int32): [0] iterations used, [1] number of volumes not converged to tol, [2..3] scratch.
 * ------------------------------------------------------------------------------------------ */
size_t leida_eigvec_dense_workspace_bytes(int N, int n_vol);
int    leida_eigvec_dense_f64(const double* dfc, int N, int n_vol, double* out, int64_t out_pitch,
                              double tol, int max_iter, int32_t* info,
                              void* workspace, size_t workspace_bytes, leida_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Stage 3 -- KMeansLeida.fit        (clustering/_clustering.py:82-153,209-232)
 *
 * A "batch" advances many independent runs (K, restart) in lock-step over the same rows.
 * The caller allocates the arrays; leida_kmeans_layout() says how big they are.
 *
 * Row data     X      fp32 (M, x_pitch), x_pitch multiple of 32, columns N..x_pitch-1 zero.
 *              xnorm  fp64 (M)  row 2-norms  (leida_kmeans_row_norms).
 * Per run r    K[r] clusters; its centroid rows are crow0[r] .. crow0[r]+K[r]-1 of the pools.
 * Pools        cent   fp32 (CR, x_pitch)   current centroids
 *              cnorm  fp64 (CR)            their 2-norms
 *              acc    int64 (CR, acc_pitch): fixed-point (2^LEIDA_KM_FRAC_BITS) column sums of the
 *                     member rows in [0,N), member count in column x_pitch   (acc_pitch = x_pitch+8)
 * Labels       int32 (R, M)   (-1 = unassigned)
 * Run state    int64 (R, LEIDA_KM_STATE_WORDS): see LEIDA_KM_ST_*
 *
 * Everything that is reduced across rows is an int64 sum, so results do not depend on the
 * order of execution nor on how rows are sharded over GPUs: with G ranks, sum `acc` and the
 * reducible state words over ranks (ncclAllReduce, sum, int64) between leida_kmeans_assign and
 * leida_kmeans_update.
 * ------------------------------------------------------------------------------------------ */
#define LEIDA_KM_FRAC_BITS     40
#define LEIDA_KM_STATE_WORDS   16
#define LEIDA_KM_ST_CHANGED    0   /* rows whose label changed in the last pass     (reducible) */
#define LEIDA_KM_ST_DIST_HI    1   /* distortion, 2^-40 units                        (reducible) */
#define LEIDA_KM_ST_DIST_LO    2   /* distortion remainder, 2^-80 units              (reducible) */
#define LEIDA_KM_ST_RECHECKED  3   /* rows re-evaluated in fp64 in the last pass     (reducible) */
#define LEIDA_KM_ST_ACTIVE     4   /* 1 while the run is still iterating                          */
#define LEIDA_KM_ST_PASSES     5   /* assignment passes done (initial pass included)              */
#define LEIDA_KM_ST_EMPTY      6   /* 1 + index of the first empty cluster seen, 0 = none         */
#define LEIDA_KM_ST_K          7   /* K of the run (set by leida_kmeans_init)                     */
#define LEIDA_KM_ST_RECHECK_TOTAL 8  /* rows re-evaluated in fp64 over all passes (global count)    */
#define LEIDA_KM_ST_CHANGED_TOTAL 9  /* label changes over all passes (global count)                */
#define LEIDA_KM_ST_TAU_WATCH  10  /* float bits: largest tensor-core margin, as a fraction of the trust threshold, at
                                      which the fp64 re-check overruled the tensor-core argmax (0 = never); local     */

#define LEIDA_KM_MAX_K         64

int leida_kmeans_row_norms(const float* X, int64_t M, int N, int64_t x_pitch, double* xnorm,
                           leida_stream_t stream);

/* Pack an arbitrary fp32 (M, N) matrix (pitch src_pitch) into the padded layout above. */
int leida_kmeans_pack_rows(const float* src, int64_t M, int N, int64_t src_pitch,
                           float* X, int64_t x_pitch, leida_stream_t stream);

/* Start the runs: centroids of run r = rows init_rows[irow0 .. ] of X_init (clustering/_clustering.py:105-108).
 * X_init/xi_pitch may be a different array from the rank's shard (multi-GPU: the gathered seed rows);
 * init_rows (device int64, CR entries) index X_init.  Clears acc, labels := -1, state. */
int leida_kmeans_init(const float* X_init, int64_t xi_pitch, const int64_t* init_rows,
                      int n_runs, const int32_t* run_k, const int32_t* run_crow0, int N, int64_t x_pitch,
                      float* cent, double* cnorm, int64_t* acc, int32_t* labels, int64_t M,
                      int64_t* state, leida_stream_t stream);

/* One assignment pass of every ACTIVE run (clustering/_clustering.py:111-114,127-128): label =
 * argmin_k cosine distance (first minimum), fp32 dot products with an fp64 re-evaluation of every
 * row whose top-2 margin is inside the fp32 error bound, so labels equal the fp64 argmin.  Rows
 * whose label changed move their fixed-point contribution between clusters in `acc`.
 * kmax (host value) = max K over the runs of this call; it selects the register tile, so callers
 * group runs of similar K.  worklist: device int32 (8*worklist_cap + 8), 16-byte aligned, header (8 words)
 * zeroed once by the caller: rows the fp32 pass cannot decide are queued there as 8-word entries
 * {run, row, fast argmax or -1, margin / threshold (float bits), candidate mask lo, hi, 0, 0} (rows that do
 * not fit are resolved inline, slower but identical) and MUST be resolved by leida_kmeans_recheck on the
 * same stream before the next leida_kmeans_assign / leida_kmeans_update. */
int leida_kmeans_assign(const float* X, const double* xnorm, int64_t M, int N, int64_t x_pitch,
                        int n_runs, int kmax, const int32_t* run_k, const int32_t* run_crow0,
                        const float* cent, const double* cnorm, int64_t* acc,
                        int32_t* labels, int64_t* state,
                        int32_t* worklist, int64_t worklist_cap, int metric, leida_stream_t stream);
/* metric: 0 = cosine distance (the LEiDA path), 1 = Euclidean distance -- KMeansLeida(metric='euclidean'),
 * clustering/_clustering.py:61-75,111,127 (cdist(..., 'euclidean')): score = x.c - |c|^2/2, exact re-evaluation
 * sqrt(sum (x - c)^2).  The tensor-core flow below is cosine only. */

/* fp64 re-evaluation of the queued rows with the reference's arithmetic (scipy cdist 'cosine':
 * dot/(|u||v|), clip, 1-cos; first minimum) over the entry's candidate clusters (empty mask = all K: the
 * candidates are the clusters whose fast score was within the trust threshold of the best, which always
 * include the fp64 argmin); empties the queue.  When the fp64 argmin differs from the entry's fast argmax the
 * entry's margin fraction is max-ed into state[LEIDA_KM_ST_TAU_WATCH]. */
int leida_kmeans_recheck(const float* X, const double* xnorm, int64_t M, int N, int64_t x_pitch,
                         int n_runs, const int32_t* run_k, const int32_t* run_crow0,
                         const float* cent, const double* cnorm, int64_t* acc,
                         int32_t* labels, int64_t* state,
                         int32_t* worklist, int64_t worklist_cap, int32_t* labels_out, int metric,
                         leida_stream_t stream);
/* labels_out == NULL: the exact label replaces labels[run][row] and the row's contribution is moved
 * in acc (SIMT flow).  labels_out != NULL: the exact label is written to labels_out[run][row] only
 * (tensor-core flow; leida_kmeans_tc_apply does the moves).  In that flow a full queue is not an error:
 * leida_kmeans_tc_assign marks every undecided pair with label -2, and when the queue overflowed the
 * re-check also scans labels_out of the n_runs runs for marks that are not in the queue. */

/* ---- tensor-core assignment flow (tcgen05 / TMEM / TMA) -------------------------------------
 * Same semantics as leida_kmeans_assign for ALL active runs in one launch:
 *   S = X * C'^T on the tensor cores with bf16 x 3 split operands (fp32-grade products), fused
 *   top-2 epilogue, fp64 re-check of every row inside the error bound.  Per pass:
 *     leida_kmeans_tc_assign  -> labels_new (label, or -2 = undecided at this precision: queued for the re-check,
 *                                or merely marked when the queue is full)
 *     leida_kmeans_recheck(labels_out = labels_new)
 *     leida_kmeans_tc_apply   -> changed counts, fixed-point moves in acc, labels_cur := labels_new
 *     [all-reduce acc/state]  leida_kmeans_update
 *     leida_kmeans_tc_plan    -> packs the still-active runs' unit-normalised centroids for the next pass
 * X1/X2: bf16 (M, tc_pitch) split of X (leida_kmeans_tc_split_rows, once per data set).
 * C1/C2: bf16 (c_rows, tc_pitch) packed centroid operands; every run owns a slot of 4..64 columns (a
 *        multiple of 4: its K unit centroids + padding columns that score -1e30 through a constant bias
 *        feature), a 128-column tile holds slots of one width; the plan may give a run a wider slot than
 *        its K needs so that part-filled tiles of different widths merge.  c_rows >= 128 * (number of
 *        tiles), a bound is 2*sum(K) + 128*12.  run_pcol[r] = first packed column | slot width << 24, or -1.  plan (2), tile_first/tile_count (c_rows/128), prun/pcol/run_pcol
 *        (n_runs): int32 device arrays owned by the caller, written by leida_kmeans_tc_plan. */
int64_t leida_kmeans_tc_pitch(int N);
/* packed centroid columns per tile of this build (128 or 256); c_rows must be a multiple, a bound on the packed columns
 * is 4*sum(K) + 17 tiles */
int leida_kmeans_tc_tile_columns(void);
int leida_kmeans_tc_split_rows(const float* X, int64_t M, int N, int64_t x_pitch, void* X1, void* X2,
                               leida_stream_t stream);
int leida_kmeans_tc_plan(int n_runs, const int32_t* run_k, const int32_t* run_crow0, const int64_t* state,
                         const float* cent, const double* cnorm, int N, int64_t x_pitch,
                         int32_t* plan, int32_t* tile_first, int32_t* tile_count, int32_t* prun, int32_t* pcol,
                         int32_t* run_pcol, void* C1, void* C2, int pack, leida_stream_t stream);
/* pack != 0 also writes C1/C2 from cent/cnorm (needed once after leida_kmeans_init; afterwards
 * leida_kmeans_update keeps them current).  The plan may be refreshed at any pass boundary (before
 * leida_kmeans_update); runs that stopped simply stay packed until the next refresh. */
int leida_kmeans_tc_assign(const void* X1, const void* X2, const double* xnorm, int64_t M, int N,
                           const void* C1, const void* C2, int64_t c_rows,
                           const int32_t* plan, const int32_t* tile_first, const int32_t* tile_count, const int32_t* prun,
                           int32_t* labels_new, int32_t* worklist, int64_t worklist_cap, double tau,
                           leida_stream_t stream);
/* tau: trust threshold on (best - second) / |x|; <= 0 selects leida_kmeans_tc_tau(N), the measured error bound
 * of the bf16 x 3 tensor-core score (kmeans_tc.cu: default_tau). */
double leida_kmeans_tc_tau(int N);
/* Same launch, additionally storing the raw fp32 tensor-core scores: scores (M, score_pitch >= c_rows), column =
 * packed centroid column (run_pcol).  Verification entry: the tests compare the scores with fp64 to measure the
 * error bound that leida_kmeans_tc_tau encodes. */
int leida_kmeans_tc_scores(const void* X1, const void* X2, const double* xnorm, int64_t M, int N,
                           const void* C1, const void* C2, int64_t c_rows,
                           const int32_t* plan, const int32_t* tile_first, const int32_t* tile_count, const int32_t* prun,
                           int32_t* labels_new, int32_t* worklist, int64_t worklist_cap,
                           float* scores, int64_t score_pitch, leida_stream_t stream);
int leida_kmeans_tc_apply(const float* X, int64_t M, int N, int64_t x_pitch, int unit_rows, int n_runs_bound,
                          const int32_t* plan, const int32_t* prun,
                          const int32_t* run_k, const int32_t* run_crow0,
                          int32_t* labels_cur, const int32_t* labels_new, int64_t* acc, int64_t* state,
                          leida_stream_t stream);

/* unit_rows != 0: the caller guarantees |X[i][j]| <= 1 for every element (true for LEiDA eigenvectors); the
 * kernel may then quantise on the FMA pipe.  Results are the same integers either way. */

/* Convergence bookkeeping + centroid update of every active run (clustering/_clustering.py:118-133):
 * a run stops when a pass at Lloyd iteration >= 1 changed no label, or after n_iter iterations;
 * otherwise its centroids become the member means (:120-125), rounded to fp32.
 * acc_reduced / state_reduced are the (all-reduced) copies to read; with one GPU pass the same
 * pointers as acc / state.  An empty cluster stops the run and sets LEIDA_KM_ST_EMPTY. */
int leida_kmeans_update(int n_runs, const int32_t* run_k, const int32_t* run_crow0, int N, int64_t x_pitch,
                        const int64_t* acc_reduced, const int64_t* state_reduced,
                        float* cent, double* cnorm, int64_t* state, int n_iter,
                        int32_t* n_active_out, const int32_t* run_pcol, void* C1, void* C2,
                        int32_t* mailbox, leida_stream_t stream);
/* Same, with the all-reduce fused in (several GPUs of one node): peer_bufs is a DEVICE array of n_peers
 * pointers, one per rank (this rank included), each to that rank's exchange buffer -- its acc words followed, from
 * word state_offset_words, by its state words -- mapped into this process (symmetric memory).  The kernel sums the
 * ranks' words itself (int64: exact, order free), only for the runs that are still active: with multicast_buf != NULL
 * (the NVSwitch multicast mapping of the same buffers) by one multimem.ld_reduce.add.u64 per word -- the switch adds --
 * else by one load per rank over NVLink.  flags/seq != NULL: the kernel waits until flags[r] >= *seq + 1 for every rank
 * r (written by the ranks' leida_kmeans_exchange_publish of this pass) and increments *seq when done; NULL: the caller
 * has ordered the ranks (every buffer complete before any rank's call starts). */
int leida_kmeans_update_peers(int n_runs, const int32_t* run_k, const int32_t* run_crow0, int N, int64_t x_pitch,
                              const int64_t* const* peer_bufs, int n_peers, const int64_t* multicast_buf,
                              int64_t state_offset_words, const uint64_t* flags, const uint64_t* seq,
                              float* cent, double* cnorm, int64_t* state, int n_iter,
                              int32_t* n_active_out, const int32_t* run_pcol, void* C1, void* C2,
                              int32_t* mailbox, leida_stream_t stream);
/* Publish this rank's words of one pass: copies the acc rows of the still-active runs and all state words from
 * local_words (acc words, then from n_acc_words the state words) to exchange_half (one half of this rank's symmetric
 * buffer, same layout), then writes ++*seq into slot my_rank of every rank's flag array (peer_flags: DEVICE array of
 * n_peers pointers) with release semantics at system scope.  Halves alternate pass by pass; seq (device uint64) and
 * counter (device int32, zero) persist across passes. */
int leida_kmeans_exchange_publish(int n_runs, const int32_t* run_k, const int32_t* run_crow0, int64_t x_pitch,
                                  const int64_t* local_words, int64_t n_acc_words, int64_t* exchange_half,
                                  uint64_t* const* peer_flags, int n_peers, int my_rank, uint64_t* seq,
                                  int32_t* counter, leida_stream_t stream);
/* The same update with the exchange fused in and the centroid rows kept in registers from the sums to the packed
 * tensor-core operands (the form the engine uses; x_pitch <= 1024).  local_words: this rank's acc words followed, from
 * word state_offset_words, by its state words (the buffer leida_kmeans_tc_apply accumulates into).
 * Single GPU: exchange_half = NULL (the other exchange arguments are ignored).
 * Several GPUs of one node: exchange_half = this rank's half (of this pass's parity) of its symmetric-memory buffer,
 * peer_halves = DEVICE array of every rank's pointer to that half, multicast_half = multicast address of the half or
 * NULL, peer_flags = DEVICE array of every rank's flag array ([n_peers][flag_pitch] uint64, zero at start),
 * flags_local = this rank's own flag array, seq = device uint64 pass counter (zero at start, same on every rank).
 * The CTA of run r publishes the run's words, stores ++pass into flag[my_rank][r] of every rank (release, system
 * scope), waits until flag[*][r] of every rank has reached it, and reads every word as the sum over ranks -- one
 * multimem.ld_reduce.add.u64 through the NVSwitch multicast mapping, or one load per rank.  Halves must alternate
 * from pass to pass.  No separate publish call, no collective, no host involvement. */
int leida_kmeans_update_fused(int n_runs, const int32_t* run_k, const int32_t* run_crow0, int N, int64_t x_pitch,
                              int64_t* local_words, int64_t state_offset_words, int64_t* exchange_half,
                              const int64_t* const* peer_halves, const int64_t* multicast_half,
                              uint64_t* const* peer_flags, const uint64_t* flags_local, int64_t flag_pitch,
                              int n_peers, int my_rank, uint64_t* seq, float* cent, double* cnorm,
                              int64_t* state, int n_iter, int32_t* n_active_out, const int32_t* run_pcol,
                              void* C1, void* C2, int32_t* mailbox, leida_stream_t stream);
/* n_active_out: device int32[8], zeroed once by the caller ([0],[1] scratch re-armed by the kernel,
 * [2] = runs still active after this call, [3] = calls completed, [4] = first all-stopped call).  mailbox (may be NULL): HOST-visible (pinned, mapped),
 * 8-byte aligned; the last CTA publishes ONE 64-bit word there -- bits 0-15 runs still active, bits 16-39 number
 * of update calls completed, bits 40-63 index of the first update call that left no run active (0 = not yet) --
 * so the host can follow convergence without synchronising or copying (n_runs <= 65535, calls < 2^24).
 * With several ranks the values are identical on every rank (they
 * derive from all-reduced counters), which lets all ranks issue the same number of passes. */
/* run_pcol / C1 / C2 (all NULL in the SIMT flow): the tensor-core operands of every packed run --
 * unit-normalised centroids split into bf16 hi/lo rows at the run's packed column -- are rewritten
 * here, so the next leida_kmeans_tc_assign needs no separate packing launch. */

/* Distortion of every run against its final centroids (clustering/_clustering.py:209-232), fp64
 * per row, accumulated as exact integers into state[DIST_HI/LO].  kmax (host value): the largest K of the
 * batch -- it sizes the shared-memory ring of the tiled kernels (rows in registers, every run's centroids staged
 * in shared memory by one bulk copy, a ring of up to six runs handed over through mbarriers; x_pitch <= 512 and a
 * multiple of 4); 0 selects the gather kernel. */
int leida_kmeans_distortion(const float* X, const double* xnorm, int64_t M, int N, int64_t x_pitch,
                            int n_runs, int kmax, const int32_t* run_k, const int32_t* run_crow0,
                            const float* cent, const double* cnorm, const int32_t* labels,
                            int64_t* state, int metric, leida_stream_t stream);

/* Reference-faithful centroid update (validation mode, single GPU): float32 row-after-row sums in
 * row order and a float32 divide, bit-identical to numpy's y[labels==c].mean(axis=0)
 * (clustering/_clustering.py:122).  Overwrites cent/cnorm of the runs whose state says active. */
int leida_kmeans_means_f32_sequential(const float* X, int64_t M, int N, int64_t x_pitch,
                                      int n_runs, const int32_t* run_k, const int32_t* run_crow0,
                                      const int32_t* labels, const int64_t* state,
                                      float* cent, double* cnorm, leida_stream_t stream);

/* KMeansLeida.transform / predict (clustering/_clustering.py:162-207): fp64 cosine distances of
 * rows to K centroids -> dist (M, K) (may be NULL) and/or labels (M) (may be NULL); metric 0 = cosine, 1 = Euclidean. */
int leida_kmeans_transform_f64(const float* X, int64_t M, int N, int64_t x_pitch,
                               const float* cent, int K, int64_t c_pitch,
                               double* dist, int32_t* labels, int metric, leida_stream_t stream);

/* labels[i] = lut[labels[i]]  (KMeansLeida._remap_labels, clustering/_clustering.py:234-247;
 * the permutation itself is computed on the host with the reference's numpy calls). */
int leida_remap_labels_i32(int32_t* labels, int64_t M, const int32_t* lut, int K, leida_stream_t stream);
/* The same for the selected run of every K at once: out[c][i] = lut[c * kmax + labels[src_run[c]][i]], labels being the
 * (n_runs, M) label array of the sweep and src_run[c] the row of the best restart of the c-th K (:139-146). */
int leida_remap_labels_batch_i32(const int32_t* labels, int64_t M, int n_out, const int32_t* src_run,
                                 const int32_t* lut, int kmax, int32_t* out, leida_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Clustering quality scores of identify_states (clustering/_clustering.py:331-339, dunn_fast :934-973;
 * sklearn silhouette_score(metric='cosine'), davies_bouldin_score)  -- SURVEY 8f-1.
 * All label columns (one per K) are scored in one call.  labels: int32 (n_cols, label_pitch) over ALL
 * rows; sel (device int64, n entries, may be NULL = the first n rows) picks the scored sub-sample
 * (the reference scores 30 000 random rows when M > 30 000, :318-320).  col0[c] = first cluster slot
 * of column c (slots are col_k[c] wide, total_clusters in all).
 * leida_scores_cosine: xs float (n, x_pitch) scratch; G int64 (total_clusters, x_pitch) and cnt
 *   int32 (total_clusters) receive the fixed-point (2^40) sums of the normalised rows per cluster and
 *   the cluster sizes; outputs per column: sil_sum int64 = sum_i silhouette_i * 2^40,
 *   min_inter / max_intra uint32 = float bits of the smallest non-zero inter-cluster and the largest
 *   intra-cluster cosine distance (Dunn = min_inter / max_intra).
 * leida_scores_db: centroids double (total_clusters, N); dist_sum int64 (total_clusters) = sum over the
 *   members of the Euclidean distance to their centroid * 2^40.
 * ------------------------------------------------------------------------------------------ */
int leida_scores_cosine(const float* X, int64_t x_pitch, int N, const int64_t* sel, int64_t n,
                        const int32_t* labels, int64_t label_pitch, int n_cols, const int32_t* col0,
                        const int32_t* col_k, int total_clusters, float* xs, int64_t* G, int32_t* cnt,
                        int64_t* sil_sum, uint32_t* min_inter, uint32_t* max_intra, leida_stream_t stream);
int leida_scores_db(const float* X, int64_t x_pitch, int N, int64_t M, const int32_t* labels,
                    int64_t label_pitch, int n_cols, const int32_t* col0, const double* centroids,
                    int total_clusters, int64_t* dist_sum, leida_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Stage 4 -- dynamics metrics counts
 *   (dynamics_metrics/_dynamics_metrics.py:131-134 occupancy, :447-453 runs, :242-245 transitions)
 * labels: int32 (n_cols, M) with pitch label_pitch; column c uses K = col_k[c] states.
 * Segment g (one (condition, subject) group) = rows seg_rows[seg_off[g] .. seg_off[g+1]) when
 * seg_rows != NULL, else the contiguous rows [seg_off[g], seg_off[g+1]).
 * Outputs int32, zero-filled by the call: occ (n_cols, G, kmax), runs (n_cols, G, kmax),
 * trans (n_cols, G, kmax, kmax).
 * ------------------------------------------------------------------------------------------ */
int leida_dynamics_counts_i32(const int32_t* labels, int64_t label_pitch, int n_cols, const int32_t* col_k,
                              const int64_t* seg_off, const int64_t* seg_rows, int n_segments, int kmax,
                              int32_t* occ, int32_t* runs, int32_t* trans, leida_stream_t stream);

/* ---- post-hoc statistics (reference: pyleida/stats/_stats.py:54-269) ------------------------
 * The two-group tests the reference runs on every column of the occupancy / dwell-time tables after the hot
 * path (permtest_ind, permtest_rel, hedges_g; _compute_stats :275-364, called from _leida.py:336-342),
 * batched over variables x permutations.  values: device fp64 (n_rows, n_vars) row-major table; idx1 / idx2:
 * device int32 row indices of the two groups.
 *
 * leida_stats_moments_f64: out[var][8] = {mean1, mean2, sum of squared deviations 1, 2, Levene W
 *   (scipy.stats.levene center='mean', :102), n1, n2, 0}; hedges_g (:135-192) and the Levene p-value follow
 *   from these on the host.
 * leida_perm_ttest_ind_f64: t_obs[var] = scipy.stats.ttest_ind statistic (pooled when equal_var[var], else
 *   Welch); counts[var][3] = number of permuted splits with t <= t_obs, t >= t_obs, |t| >= |t_obs|
 *   (the comparisons of SciPy's _permutation_ttest, which ttest_ind(permutations=) ran, :105-111).
 *   perms == NULL: n_perm uniformly random splits drawn on the device (Philox4x32-10 keyed by seed, selection
 *   sampling).  perms != NULL: int32 (n_perm, n1+n2) index permutations of the pooled sample (group 1 first),
 *   e.g. all C(n1+n2, n1) splits for an exact test.  null_out (may be NULL): fp64 (n_vars, n_perm).
 * leida_perm_paired_f64: obs[var] = mean(x1) - mean(x2) over the n pairs (idx1[i], idx2[i]);
 *   counts[var][2] = number of permutations with statistic <= obs + g, >= obs - g, g = 100 eps |obs|
 *   (scipy.stats.permutation_test permutation_type='samples', :236-244).  swaps == NULL: every pair swapped
 *   with probability 1/2 on the device; swaps != NULL: uint8 (n_perm, n), 1 = swap.
 * Limits: n1, n2 >= 2 (independent), n1 + n2 <= 4096, n <= 4096. */
int leida_stats_moments_f64(const double* values, int64_t n_rows, int n_vars, const int32_t* idx1, int n1,
                            const int32_t* idx2, int n2, double* out, leida_stream_t stream);
int leida_perm_ttest_ind_f64(const double* values, int64_t n_rows, int n_vars, const int32_t* idx1, int n1,
                             const int32_t* idx2, int n2, const uint8_t* equal_var, int n_perm,
                             const int32_t* perms, uint64_t seed, double* t_obs, int32_t* counts,
                             double* null_out, leida_stream_t stream);
int leida_perm_paired_f64(const double* values, int64_t n_rows, int n_vars, const int32_t* idx1,
                          const int32_t* idx2, int n, int n_perm, const uint8_t* swaps, uint64_t seed,
                          double* obs, int32_t* counts, double* null_out, leida_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LEIDA_B200_H */
