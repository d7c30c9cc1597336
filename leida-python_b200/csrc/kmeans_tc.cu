// Tensor-core assignment pass of the lock-step k-means (sm_100a: TMA + tcgen05.mma + TMEM).
//
// The distance step of every active run is one contraction  S = X (M x N) * C'^T (N x sum K),
// C' = unit-normalised centroids of all active runs packed side by side (clustering/_clustering.py:127
// evaluated for every run at once).  fp32 operands are split into bf16 pairs (v = v1 + v2 + r,
// |r| <= 2^-18 |v|) and the product is accumulated as X1*C1 + X1*C2 + X2*C1 in fp32 in TMEM, which
// reproduces the fp32 product to ~2^-16 relative; the epilogue takes the top-2 scores per (row, run),
// trusts the argmax only when the margin exceeds the error bound (default_tau below: rigorous terms + the MEASURED
// accumulation error of the tensor core, watched in production) and queues the other pairs -- with the set of
// clusters that can still be the fp64 argmin -- for the fp64 re-check (leida_kmeans_recheck), so labels are still the
// fp64 argmin of the reference.
//
// CTA = 18 warps: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM owner), warps 2..17 epilogue
// (four per TMEM lane quarter).  Persistent over row tiles of 128 rows; for every row tile all column
// tiles (128 packed centroid columns) are visited, K streamed in 64-element (128-byte, swizzled) blocks.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>
#include <mutex>

namespace leida {

constexpr int kStateWordsTc = LEIDA_KM_STATE_WORDS;
constexpr int BM = 128;          // rows per tile  (UMMA M)
#ifndef LEIDA_TC_BN
#define LEIDA_TC_BN 128
#endif
#ifndef LEIDA_MBAR_SUSPEND_NS
#define LEIDA_MBAR_SUSPEND_NS 4000      // 0: plain polling (measured: assign 595 -> 573 ms per config-3 step with the hint)
#endif
constexpr int BN = LEIDA_TC_BN;  // packed centroid columns per tile (UMMA N): 128, or 256 (one MMA does twice the work)
constexpr int BK = 64;           // bf16 elements per k-block = one 128-byte swizzle row
constexpr int ST = BN == 128 ? 3 : 2;                 // smem pipeline stages
constexpr int A_BYTES = BM * BK * 2;                  // one bf16 block of the row tile: 16 KB
constexpr int B_BYTES = BN * BK * 2;                  // one bf16 block of the column tile: 16 / 32 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // X1, X2, C1, C2 blocks: 64 / 96 KB
constexpr int EPI_SPLIT = 4;                          // epilogue warps per TMEM lane quarter
constexpr int EPI_WARPS = 4 * EPI_SPLIT;
constexpr int TC_THREADS = 64 + 32 * EPI_WARPS;
constexpr int N_ACC = 512 / BN;                       // fp32 accumulators of BN columns in flight (all 512 TMEM columns)
constexpr int TMEM_COLS = N_ACC * BN;
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((BN >> 3) << 17) | ((BM >> 4) << 24);

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    long long t0 = 0;
    for (unsigned spin = 0; !done; ++spin) {
#if LEIDA_MBAR_SUSPEND_NS > 0
        // suspend-time hint: the waiting warp sleeps in hardware until the phase completes (or the hint expires) instead
        // of polling -- 16 epilogue warps poll 2e8 times per full-width launch otherwise
        asm volatile(
            "{\n.reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
            "selp.b32 %0, 1, 0, P1;\n}\n"
            : "=r"(done) : "r"(addr), "r"(parity), "r"((uint32_t)LEIDA_MBAR_SUSPEND_NS) : "memory");
#else
        asm volatile(
            "{\n.reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.b32 %0, 1, 0, P1;\n}\n"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
#endif
        // A lost arrival must fail loudly, not hang the GPU -- but by TIME, not by poll count: preemption, MPS
        // time-slicing or a debugger can stretch a legitimate wait arbitrarily in polls.  ~20 s of SM clock.
#if LEIDA_MBAR_SUSPEND_NS > 0
        if (!done && (spin & 0x3FFu) == 0x3FFu) {
#else
        if (!done && (spin & 0xFFFFFu) == 0xFFFFFu) {
#endif
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > (40ll << 30)) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
// K-major, 128-byte swizzle, dense 8-row groups (1024 B): see cute::UMMA::SmemDescriptor.
__device__ __forceinline__ uint64_t umma_desc(const void* tile) {
    const uint64_t addr = (uint64_t)(smem_u32(tile) & 0x3FFFF) >> 4;
    return addr | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x64(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int W> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* r);
template <> __device__ __forceinline__ void tmem_ld<4>(uint32_t t, uint32_t* r) { tmem_ld_x4(t, r); }
template <> __device__ __forceinline__ void tmem_ld<8>(uint32_t t, uint32_t* r) { tmem_ld_x8(t, r); }
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t t, uint32_t* r) { tmem_ld_x16(t, r); }
template <> __device__ __forceinline__ void tmem_ld<32>(uint32_t t, uint32_t* r) { tmem_ld_x32(t, r); }
template <> __device__ __forceinline__ void tmem_ld<64>(uint32_t t, uint32_t* r) { tmem_ld_x64(t, r); }
template <> __device__ __forceinline__ void tmem_ld<0>(uint32_t, uint32_t*) {}

__host__ __device__ __forceinline__ int read_width(int K) { return tc_read_width(K); }

// Top-2 of a run's scores held in registers (read from TMEM).  The run owns a slot of W columns: its
// K centroids followed by W-K padding columns that score -1e30 by construction (bias feature, see
// k_split_rows / pack), so no per-column masking is needed.  The cluster index is packed into the 6 low
// mantissa bits of each score (a 2^-17 relative perturbation, covered by the error bound), which makes
// every key unique.  (best, second) are computed together in one tree of pair merges:
//   (m1,s1) + (m2,s2) -> (max(m1,m2), max3(min(m1,m2), s1, s2))       ~2.2 W min/max operations in all.
// three-input maximum: one FMNMX3 on sm_100
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int W>
__device__ __forceinline__ void top2_regs(const uint32_t* r, float& best, float& second, int& bk) {
    constexpr int P = W / 2;
    float m[P], s2[P];
#pragma unroll
    for (int k = 0; k < P; ++k) {
        const float a = __uint_as_float((r[2 * k] & 0xFFFFFFC0u) | (uint32_t)(2 * k));
        const float b = __uint_as_float((r[2 * k + 1] & 0xFFFFFFC0u) | (uint32_t)(2 * k + 1));
        m[k] = fmaxf(a, b);
        s2[k] = fminf(a, b);
    }
#pragma unroll
    for (int o = 1; o < P; o <<= 1)
#pragma unroll
        for (int k = 0; k + o < P; k += 2 * o) {
            const float lo = fminf(m[k], m[k + o]);
            m[k] = fmaxf(m[k], m[k + o]);
            s2[k] = fmax3(lo, s2[k], s2[k + o]);
        }
    bk = (int)(__float_as_uint(m[0]) & 63u);
    best = __uint_as_float(__float_as_uint(m[0]) & 0xFFFFFFC0u);
    second = __uint_as_float(__float_as_uint(s2[0]) & 0xFFFFFFC0u);
}

// The registers written by tcgen05.ld become valid only after tcgen05.wait::ld; this empty asm makes
// every later use of them depend on a point after the wait.
template <int W>
__device__ __forceinline__ void reg_fence(uint32_t* r) {
#pragma unroll
    for (int k = 0; k < W; ++k) asm volatile("" : "+r"(r[k]));
}

// Packing plan written by k_plan (device), read by the assign kernel.
//   plan[0] = number of column tiles, plan[1] = number of packed (active) runs
//   tile_first[t] & 0xffff, tile_count[t] : slice of the packed run list that lives in tile t;
//   tile_first[t] >> 16 = TMEM read width of the tile's widest run
//   prun[i] = run id (the runs of tile t sit at columns 0, W, 2W, ... in this order);
//   pcol[i] = (first column of the run inside its tile) | (K << 8) is written by k_plan for callers, not read here
struct TcParams {
    long long M;
    int nkb;   // k-blocks = NPK / 64
    const int32_t* plan;
    const int32_t* tile_first;
    const int32_t* tile_count;
    const int32_t* prun;
    const double* xnorm;
    float tau;
    int32_t* labels_new;
    int32_t* worklist;
    long long worklist_cap;
    float* dump;             // DUMP instantiation only: raw fp32 scores, (M, dump_pitch), column = packed column
    long long dump_pitch;
};

// Columns one epilogue step reads from TMEM for a tile whose slots are W wide: a whole number of slots,
// a divisor-friendly size so that no step reaches past the tile's BN columns, and (for the narrow
// slots that dominate the sweep) as wide as one tcgen05.ld allows, because the epilogue is bound by the
// TMEM load round trip, not by the arithmetic.
__host__ __device__ constexpr int chunk_cols(int W) { return W == 4 || W == 8 || W == 16 ? 32 : W == 12 ? 24 : W; }

template <int CH>
__device__ __forceinline__ void tmem_ld_chunk(uint32_t t, uint32_t* r) {
    if constexpr (CH == 64) tmem_ld_x64(t, r);
    else {
        constexpr int A = CH >= 32 ? 32 : 16;      // CH in {20, 24, 32, 40, 48}
        tmem_ld<A>(t, r);
        tmem_ld<CH - A>(t + A, r + A);
    }
}

// W columns, W a multiple of 4 up to 64: largest power-of-two piece first.
template <int W>
__device__ __forceinline__ void tmem_ld_any(uint32_t t, uint32_t* r) {
    if constexpr (W > 0) {
        constexpr int A = W >= 64 ? 64 : W >= 32 ? 32 : W >= 16 ? 16 : W >= 8 ? 8 : 4;
        tmem_ld<A>(t, r);
        tmem_ld_any<W - A>(t + A, r + A);
    }
}

// Epilogue of one column tile for one warp.  All slots of a tile have the same width W and sit at
// columns 0, W, 2W, ... (k_plan), so the warp reads the tile in chunks of CH columns = CH/W runs per
// tcgen05.ld and evaluates the runs from registers; the EPI_SPLIT warps of a lane quarter take the chunks
// round-robin.  (Software pipelining the loads was measured and bought nothing: with this many warps the
// scheduler already overlaps one warp's TMEM round trip with another's arithmetic.)
template <int W, bool DUMP>
__device__ __forceinline__ void epilogue_tile(const TcParams& p, uint32_t taddr, int first, int cnt, int sub, int lane,
                                              bool valid, float thr, int32_t* lab, long long row, int ct) {
    constexpr int CH = chunk_cols(W), RPC = CH / W;
    const int n_chunks = (cnt + RPC - 1) / RPC;
    if (sub >= n_chunks) return;
    const int m_run = p.prun[first + (lane < cnt ? lane : 0)];     // a tile holds at most 32 runs
    uint32_t r[CH];
    for (int c = sub; c < n_chunks; c += EPI_SPLIT) {
        tmem_ld_chunk<CH>(taddr + c * CH, r);
        tmem_ld_wait();
        reg_fence<CH>(r);
        if constexpr (DUMP) {
            if (valid) {
                float* d = p.dump + row * p.dump_pitch + (long long)ct * BN + c * CH;
#pragma unroll
                for (int i = 0; i < CH; ++i) d[i] = __uint_as_float(r[i]);
            }
        }
#pragma unroll
        for (int j = 0; j < RPC; ++j) {
            const int i = c * RPC + j;
            if (i < cnt) {                                         // warp-uniform
                const int run = __shfl_sync(0xffffffffu, m_run, i);
                float best, second;
                int bk;
                top2_regs<W>(r + j * W, best, second, bk);
                int out = bk;
                const bool undecided = valid && !((best - second) > thr);
                if (__any_sync(0xffffffffu, undecided)) {
                    // Undecided at this precision (rare: a fraction of a percent of the pairs): mark the pair and queue it
                    // for the fp64 re-check together with the set of clusters that can still be the fp64 argmin -- every
                    // score within thr of the best (the true best's tensor-core score is at most two score errors below
                    // the tensor-core best, and thr covers more than that) -- so the re-check evaluates 2-3 centroids,
                    // not K.  The slot's scores are read again from TMEM (warp-collective, hence the warp-uniform branch)
                    // instead of being kept alive in registers through the top-2 tree.  A full queue is fine: the re-check
                    // finds unlisted marks by scanning (k_recheck) and evaluates all K for them.
                    uint32_t q[W];
                    tmem_ld_any<W>(taddr + c * CH + j * W, q);
                    tmem_ld_wait();
                    reg_fence<W>(q);
                    if (undecided) {
                        out = -2;
                        const float lo = best - thr;
                        unsigned long long mask = 0;
#pragma unroll
                        for (int k = 0; k < W; ++k)
                            if (__uint_as_float(q[k] & 0xFFFFFFC0u) >= lo) mask |= 1ull << k;
                        const int slot = atomicAdd(p.worklist, 1);
                        if (slot < p.worklist_cap) {
                            // margin as a fraction of the threshold: the re-check records the largest one at which the
                            // fp64 argmin differed from the tensor-core argmax (LEIDA_KM_ST_TAU_WATCH)
                            const float ratio = thr > 0.f ? (best - second) / thr : 0.f;
                            int4* e = reinterpret_cast<int4*>(p.worklist + 8) + 2 * (long long)slot;
                            e[0] = make_int4(run, (int32_t)row, bk, __float_as_int(ratio));
                            e[1] = make_int4((int)(uint32_t)mask, (int)(uint32_t)(mask >> 32), 0, 0);
                        }
                    }
                }
                if (valid) {
                    lab[(long long)run * p.M] = out;
                }
            }
        }
    }
}

template <bool DUMP>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_assign(const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
            const __grid_constant__ CUtensorMap tmC1, const __grid_constant__ CUtensorMap tmC2,
            const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* stages = base;                                         // ST * 64 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + ST * STAGE_BYTES);
    uint64_t* full = bars;            // [ST]
    uint64_t* empty = bars + ST;      // [ST]
    uint64_t* tfull = bars + 2 * ST;      // [N_ACC]
    uint64_t* tempty = tfull + N_ACC;     // [N_ACC]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + N_ACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < N_ACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 32 * EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_ct = p.plan[0];
    const int n_rt = (int)((p.M + BM - 1) / BM);

    if (warp == 0) {
        if (lane == 0 && n_ct > 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int rt = blockIdx.x; rt < n_rt; rt += gridDim.x)
                for (int ct = 0; ct < n_ct; ++ct)
                    for (int kb = 0; kb < p.nkb; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        unsigned char* s = stages + stage * STAGE_BYTES;
                        mbar_expect_tx(&full[stage], STAGE_BYTES);
                        tma_load_2d(s, &tmX1, &full[stage], kb * BK, rt * BM);
                        tma_load_2d(s + A_BYTES, &tmX2, &full[stage], kb * BK, rt * BM);
                        tma_load_2d(s + 2 * A_BYTES, &tmC1, &full[stage], kb * BK, ct * BN);
                        tma_load_2d(s + 2 * A_BYTES + B_BYTES, &tmC2, &full[stage], kb * BK, ct * BN);
                        if (++stage == ST) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_ct > 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, aphase = 0;
            for (int rt = blockIdx.x; rt < n_rt; rt += gridDim.x)
                for (int ct = 0; ct < n_ct; ++ct) {
                    mbar_wait(&tempty[acc], aphase ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + acc * BN;
                    for (int kb = 0; kb < p.nkb; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        unsigned char* s = stages + stage * STAGE_BYTES;
                        const uint64_t a1 = umma_desc(s), a2 = umma_desc(s + A_BYTES);
                        const uint64_t b1 = umma_desc(s + 2 * A_BYTES), b2 = umma_desc(s + 2 * A_BYTES + B_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            tc_mma_bf16(d, a1 + 2 * k, b1 + 2 * k, (kb | k) ? 1u : 0u);
                            tc_mma_bf16(d, a1 + 2 * k, b2 + 2 * k, 1u);
                            tc_mma_bf16(d, a2 + 2 * k, b1 + 2 * k, 1u);
                        }
                        tc_commit(&empty[stage]);
                        if (++stage == ST) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(&tfull[acc]);
                    if (++acc == N_ACC) { acc = 0; aphase ^= 1; }
                }
        }
    } else if (n_ct > 0) {
        // Epilogue: EPI_WARPS warps, EPI_SPLIT per TMEM lane quarter, which share every column tile's chunks.
        const int q = warp & 3;                       // TMEM lane quarter this warp may touch
        const int sub = (warp - 2) >> 2;              // which of the quarter's EPI_SPLIT warps
        const int r_in_tile = q * 32 + lane;
        int acc = 0;
        uint32_t aphase = 0;
        for (int rt = blockIdx.x; rt < n_rt; rt += gridDim.x) {
            const long long row = (long long)rt * BM + r_in_tile;
            const bool valid = row < p.M;
            const float thr = valid ? p.tau * (float)p.xnorm[row] : 0.f;
            int32_t* lab = p.labels_new + row;
            for (int ct = 0; ct < n_ct; ++ct) {
                const int tf = p.tile_first[ct], cnt = p.tile_count[ct];
                const int first = tf & 0xffff, wt = tf >> 16;      // wt = slot width of the tile's runs
                mbar_wait(&tfull[acc], aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
                switch (wt) {
                    case 4: epilogue_tile<4, DUMP>(p, taddr, first, cnt, sub, lane, valid, thr, lab, row, ct); break;
                    case 8: epilogue_tile<8, DUMP>(p, taddr, first, cnt, sub, lane, valid, thr, lab, row, ct); break;
                    case 12: epilogue_tile<12, DUMP>(p, taddr, first, cnt, sub, lane, valid, thr, lab, row, ct); break;
                    case 16: epilogue_tile<16, DUMP>(p, taddr, first, cnt, sub, lane, valid, thr, lab, row, ct); break;
                    case 20: epilogue_tile<20, DUMP>(p, taddr, first, cnt, sub, lane, valid, thr, lab, row, ct); break;
                    case 24: epilogue_tile<24, DUMP>(p, taddr, first, cnt, sub, lane, valid, thr, lab, row, ct); break;
                    case 32: epilogue_tile<32, DUMP>(p, taddr, first, cnt, sub, lane, valid, thr, lab, row, ct); break;
                    case 40: epilogue_tile<40, DUMP>(p, taddr, first, cnt, sub, lane, valid, thr, lab, row, ct); break;
                    case 48: epilogue_tile<48, DUMP>(p, taddr, first, cnt, sub, lane, valid, thr, lab, row, ct); break;
                    default: epilogue_tile<64, DUMP>(p, taddr, first, cnt, sub, lane, valid, thr, lab, row, ct); break;
                }
                tc_fence_before();
                mbar_arrive(&tempty[acc]);            // accumulator may be overwritten by the next tile
                if (++acc == N_ACC) { acc = 0; aphase ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// ---- operand preparation ----------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// X fp32 (M, XP) -> X1, X2 bf16 (M, NPK); columns >= N are zero.
__global__ void k_split_rows(const float* __restrict__ X, long long M, int N, long long XP, __nv_bfloat16* __restrict__ X1,
                             __nv_bfloat16* __restrict__ X2, int NPK) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * NPK) return;
    const long long r = i / NPK;
    const int j = (int)(i - r * NPK);
    // column N is the constant bias feature (1 in the hi part): real centroids carry 0 there, the padding
    // columns of a run's slot carry -1e30, so a padding column always scores -1e30
    const float v = j < N ? X[r * XP + j] : (j == N ? 1.f : 0.f);
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    X1[i] = hi;
    X2[i] = lo;
}

// Packing plan: every active run gets a slot of W columns (W >= read_width(K)) inside a BN-column tile
// whose slots all have the same width.  Normally W = read_width(K); but the runs of a width class that
// would only part-fill a tile are moved up into a wider class when that class's own part-filled tile has
// room for them, so late passes with a handful of runs of different K need one tile, not one per class.
// One CTA: the run table is staged in shared memory with coalesced loads, thread 0 does the (inherently
// sequential, a few cycles per run) placement, all threads write the result back.
// run_pcol[r] = packed column | W << 24 (or -1); prun/pcol are ordered by (tile, column).
// Runs a tile can hold for slot width w: the epilogue indexes a tile's runs by lane (<= 32) and reads whole chunks of
// chunk_cols(w) columns, so the count is a multiple of the runs per chunk and no chunk reaches past the tile.
__host__ __device__ constexpr int tile_capacity(int w) {
    const int rpc = chunk_cols(w) / w;
    int cap = BN / w;
    if (cap > 32) cap = 32;
    return cap / rpc * rpc;
}
constexpr int PLAN_MAX_RUNS = 4096;
constexpr int PLAN_MAX_TILES = PLAN_MAX_RUNS / 2;   // widest slot = 64 columns = two runs per tile
constexpr int PLAN_CLASSES = 17;              // slot width = 4 * class, classes 1..16
__global__ void __launch_bounds__(256)
k_plan(int n_runs, const int32_t* __restrict__ run_k, const long long* __restrict__ state, int32_t* plan,
       int32_t* tile_first, int32_t* tile_count, int32_t* prun, int32_t* pcol, int32_t* run_pcol) {
    __shared__ short s_k[PLAN_MAX_RUNS];        // K, or 0 when the run is not active
    __shared__ int s_pc[PLAN_MAX_RUNS];         // packed global column | slot width << 24, or -1
    __shared__ short s_tfirst[PLAN_MAX_TILES];  // per tile: packed index of its first run
    __shared__ short s_tcount[PLAN_MAX_TILES];  // per tile: runs
    __shared__ unsigned char s_tw[PLAN_MAX_TILES];  // per tile: slot width
    __shared__ unsigned char s_cls[PLAN_MAX_RUNS];  // width class (slot width / 4) of the run, 0 when not active
    __shared__ int s_cnt[PLAN_CLASSES];
    __shared__ int s_tiles, s_np;
    if (threadIdx.x < PLAN_CLASSES) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    // everything that is per run and order free happens in parallel; thread 0 only does the ordered placement
    for (int r = threadIdx.x; r < n_runs; r += blockDim.x) {
        const int K = state[(long long)r * kStateWordsTc + LEIDA_KM_ST_ACTIVE] != 0 ? run_k[r] : 0;
        const int c = K > 0 ? read_width(K) >> 2 : 0;
        s_k[r] = (short)K;
        s_cls[r] = (unsigned char)c;
        if (c) atomicAdd(&s_cnt[c], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int cnt[PLAN_CLASSES], target[PLAN_CLASSES], stay[PLAN_CLASSES], cur_tile[PLAN_CLASSES], cur_n[PLAN_CLASSES];
        int np = 0;
        for (int c = 0; c < PLAN_CLASSES; ++c) { cnt[c] = s_cnt[c]; np += cnt[c]; target[c] = 0; cur_tile[c] = -1; cur_n[c] = 0; }
        // Runs of one width class come in long stretches (the sweep is K-major), so the state of the current class
        // lives in registers and the per-class arrays (local memory: dynamic indexing) are touched only on a change.
        // ascending: the remainder of class c moves to the narrowest wider class whose part-filled tile can take it
        for (int c = 1; c < PLAN_CLASSES; ++c) {
            const int cap = tile_capacity(4 * c), rem = cnt[c] % cap;
            if (rem == 0) continue;
            for (int d = c + 1; d < PLAN_CLASSES; ++d) {
                const int capd = tile_capacity(4 * d), remd = cnt[d] % capd;
                if (remd != 0 && remd + rem <= capd) { target[c] = d; cnt[c] -= rem; cnt[d] += rem; break; }
            }
        }
        for (int c = 0; c < PLAN_CLASSES; ++c) stay[c] = cnt[c];
        int n_tiles = 0;
        int cc = -1, c_stay = 0, c_tile = -1, c_n = 0, c_cap = 1, c_w = 4;
        auto flush = [&]() {
            if (cc >= 0) {
                stay[cc] = c_stay; cur_tile[cc] = c_tile; cur_n[cc] = c_n;
                if (c_tile >= 0) s_tcount[c_tile] = (short)c_n;
            }
        };
        auto load = [&](int c) {
            cc = c; c_stay = stay[c]; c_tile = cur_tile[c]; c_n = cur_n[c]; c_w = 4 * c; c_cap = tile_capacity(c_w);
        };
        for (int r = 0; r < n_runs; ++r) {
            const int c = s_cls[r];
            int v = -1;
            if (c) {
                if (c != cc) { flush(); load(c); }
                while (c_stay == 0 && target[cc] != 0) {           // arrivals beyond the class's share move on
                    const int d = target[cc];
                    flush();
                    load(d);
                }
                --c_stay;
                if (c_tile < 0 || c_n == c_cap) {
                    if (c_tile >= 0) s_tcount[c_tile] = (short)c_n;
                    c_tile = n_tiles++;
                    c_n = 0;
                    s_tw[c_tile] = (unsigned char)c_w;
                }
                v = (c_tile * BN + c_n * c_w) | (c_w << 24);
                ++c_n;
            }
            s_pc[r] = v;
        }
        flush();
        int first = 0;
        for (int t = 0; t < n_tiles; ++t) { s_tfirst[t] = (short)first; first += s_tcount[t]; }
        s_tiles = n_tiles;
        s_np = np;
    }
    __syncthreads();
    const int n_tiles = s_tiles;
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
        tile_first[t] = (int)s_tfirst[t] | ((int)s_tw[t] << 16);
        tile_count[t] = s_tcount[t];
    }
    for (int r = threadIdx.x; r < n_runs; r += blockDim.x) {
        const int v = s_pc[r];
        run_pcol[r] = v;
        if (v >= 0) {
            const int w = v >> 24, pc = v & 0xFFFFFF, t = pc / BN, col = pc - t * BN;
            const int idx = s_tfirst[t] + col / w;
            prun[idx] = r;
            pcol[idx] = col | ((int)s_k[r] << 8);
        }
    }
    if (threadIdx.x == 0) { plan[0] = n_tiles; plan[1] = s_np; }
}

// Unit-normalised, bf16-split centroid rows of the active runs at their packed position.
__global__ void k_pack_centroids(const int32_t* __restrict__ run_k, const int32_t* __restrict__ run_crow0,
                                 const int32_t* __restrict__ run_pcol, const float* __restrict__ cent,
                                 const double* __restrict__ cnorm, int N, int XP, __nv_bfloat16* __restrict__ C1,
                                 __nv_bfloat16* __restrict__ C2, int NPK) {
    const int run = blockIdx.x;
    const int v = run_pcol[run];
    if (v < 0) return;
    const int K = run_k[run], crow0 = run_crow0[run];
    const int W = v >> 24, pc = v & 0xFFFFFF;            // slot assigned by k_plan (>= read_width(K))
    for (int i = threadIdx.x; i < W * NPK; i += blockDim.x) {
        const int k = i / NPK, j = i - k * NPK;
        float v = 0.f;
        if (k >= K) {
            v = j == N ? -1e30f : 0.f;                  // padding column of the slot: scores -1e30 for every row
        } else if (j < N) {
            const double cn = cnorm[crow0 + k];
            const float inv = cn > 0.0 ? (float)(1.0 / cn) : 0.f;
            v = cent[(long long)(crow0 + k) * XP + j] * inv;
        }
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        C1[(long long)(pc + k) * NPK + j] = hi;
        C2[(long long)(pc + k) * NPK + j] = lo;
    }
}

// labels_cur vs labels_new of the packed runs: count the changes, move the fixed-point row
// contributions between clusters, and make the new labels current.
// One CTA = (chunk of APPLY_ROWS rows, packed run).  The chunk's changes are counting-sorted by
// cluster in shared memory; then thread (g, j) sums column j over its share of the rows entering
// (leaving) cluster k in a REGISTER -- independent loads, no read-modify-write chain -- and issues one
// global int64 atomic per (cluster, column) it touched.  Integer sums: any grouping gives the same result.
constexpr int APPLY_ROWS = 4096;
constexpr int APPLY_THREADS = 256;
constexpr int APPLY_DIRECT_MAX = 32;      // changed rows per (chunk, run) up to which the unsorted path is used

__device__ __forceinline__ long long quant40(float v) { return __double2ll_rn((double)v * 1099511627776.0); }

// Sum over the listed rows of quant40(x[row]) = rne(x * 2^40).  ncu (profiles/r01_k_diff_apply_ncu.txt): in the
// dense early passes the kernel is bound by instruction issue and by the conversion pipe (F2F + F2I.S64 per
// element, 16 lanes/clk/SM), not by memory.  So: the gather address is a 32-bit product (local row < 4096,
// pitch < 2^19) on one 64-bit base, eight loads are issued before the first value is used, and -- when the
// caller guarantees |x| <= 1 (UNIT: LEiDA eigenvectors are unit vectors) -- half of the elements are quantised
// on the FMA pipe instead: with h = rne(x * 2^20) and r = x - h * 2^-20 (exact, |r| <= 2^-21),
// x * 2^40 = h * 2^20 + r * 2^40 exactly and the first term is an even integer, so
// rne(x * 2^40) = h * 2^20 + rne(r * 2^40); both roundings are one fp32 fma onto 1.5 * 2^23, whose bit pattern
// is 0x4B400000 + the integer.  The bit patterns are summed as they are in 32-bit wrap-around arithmetic and
// the offsets removed every 512 elements (|sum h| <= 2^29).  Same integers as quant40 (checked exhaustively
// enough on the host: 2e8 random bit patterns and the tie / denormal / +-1 edge cases).
#ifndef LEIDA_APPLY_GATHERS
#define LEIDA_APPLY_GATHERS 8
#endif
// Four adjacent columns per thread (one 16-byte gather per listed row: a quarter of the row-index reads, address
// products and load instructions of a column-per-thread loop -- the kernel is bound by instruction issue).
// out[c] += sum over the listed rows of quant40(x[row][j4 + c]); when UNIT every element takes the FMA-pipe path.
template <bool UNIT>
__device__ __forceinline__ void quant_sum4(const float* __restrict__ xb, unsigned j4, unsigned XP,
                                           const unsigned short* __restrict__ rows, int e0, int e1, int stride,
                                           long long* out) {
    constexpr float MAGIC = 12582912.f;           // 1.5 * 2^23
    constexpr uint32_t MAGIC_BITS = 0x4B400000u;
    auto fma_path = [&](float x, uint32_t& sh, uint32_t& sl) {
        const float th = fmaf(x, 1048576.f, MAGIC);                 // MAGIC + h
        const float r = fmaf(th - MAGIC, -9.5367431640625e-07f, x); // x - h * 2^-20
        const float tl = fmaf(r, 1099511627776.f, MAGIC);           // MAGIC + rne(r * 2^40)
        sh += __float_as_uint(th);
        sl += __float_as_uint(tl);
    };
    int e = e0;
    constexpr int G = LEIDA_APPLY_GATHERS;      // 16-byte gathers in flight per thread before the first value is used
    while (e + (G - 1) * stride < e1) {
        uint32_t sh[4] = {0, 0, 0, 0}, sl[4] = {0, 0, 0, 0}, nb = 0;
        for (; nb < 512 / G && e + (G - 1) * stride < e1; ++nb, e += G * stride) {
            float4 v[G];
#pragma unroll
            for (int u = 0; u < G; ++u)
                v[u] = __ldg(reinterpret_cast<const float4*>(xb + ((unsigned)rows[e + u * stride] * XP + j4)));
#pragma unroll
            for (int u = 0; u < G; ++u) {
                if (UNIT) {
                    // all four on the FMA pipe: with 16-byte gathers the conversion pipe (F2F + F2I.S64, 16 lanes per
                    // clock and SM) was what bound the dense passes when half of the elements still went through it
                    fma_path(v[u].x, sh[0], sl[0]);
                    fma_path(v[u].y, sh[1], sl[1]);
                    fma_path(v[u].z, sh[2], sl[2]);
                    fma_path(v[u].w, sh[3], sl[3]);
                } else {
                    out[0] += quant40(v[u].x);
                    out[1] += quant40(v[u].y);
                    out[2] += quant40(v[u].z);
                    out[3] += quant40(v[u].w);
                }
            }
        }
        if (UNIT) {          // G * nb <= 512 elements per column were summed with the offset in place (|sum h| <= 2^29)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                out[c] += ((long long)(int)(sh[c] - (uint32_t)G * nb * MAGIC_BITS) << 20) +
                          (long long)(int)(sl[c] - (uint32_t)G * nb * MAGIC_BITS);
        }
    }
    for (; e < e1; e += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xb + ((unsigned)rows[e] * XP + j4)));
        out[0] += quant40(v.x); out[1] += quant40(v.y); out[2] += quant40(v.z); out[3] += quant40(v.w);
    }
}

template <bool UNIT>
__global__ void __launch_bounds__(APPLY_THREADS, 3)
k_diff_apply(const float* __restrict__ X, long long M, int N, int XP, const int32_t* __restrict__ plan,
             const int32_t* __restrict__ prun, const int32_t* __restrict__ run_k,
             const int32_t* __restrict__ run_crow0, int32_t* __restrict__ labels_cur,
             const int32_t* __restrict__ labels_new, long long* acc, long long* state) {
    __shared__ int s_list[APPLY_ROWS];       // local row | old<<16 | new<<24
    __shared__ unsigned short s_in[APPLY_ROWS];       // local rows sorted by NEW cluster
    __shared__ unsigned short s_out[APPLY_ROWS];      // local rows sorted by OLD cluster (old >= 0 only)
    __shared__ int s_cin[LEIDA_KM_MAX_K + 1], s_cout[LEIDA_KM_MAX_K + 1], s_pin[LEIDA_KM_MAX_K], s_pout[LEIDA_KM_MAX_K];
    __shared__ int s_n;
    const int AP = XP + 8;
    // blockIdx.y = row chunk, blockIdx.x = run slot: the CTAs that share a chunk have consecutive ids, are
    // scheduled together and find the chunk's rows in L2 after the first of them has read them
    const long long row_base = (long long)blockIdx.y * APPLY_ROWS;
    const int n_packed = plan[1];
  for (int pr = blockIdx.x; pr < n_packed; pr += gridDim.x) {      // persistent over the packed runs
    const int run = prun[pr];
    if (state[(long long)run * kStateWordsTc + LEIDA_KM_ST_ACTIVE] == 0) continue;
    const int K = run_k[run], crow0 = run_crow0[run];
    __syncthreads();
    int32_t* cur = labels_cur + (long long)run * M;
    const int32_t* nw = labels_new + (long long)run * M;
    if (threadIdx.x == 0) s_n = 0;
    if (threadIdx.x <= LEIDA_KM_MAX_K) { s_cin[threadIdx.x] = 0; s_cout[threadIdx.x] = 0; }
    __syncthreads();
    {   // all label loads of the chunk first (independent, coalesced), then the compares
        constexpr int PER = APPLY_ROWS / APPLY_THREADS;
        int lo[PER], ln[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const long long row = row_base + threadIdx.x + i * APPLY_THREADS;
            const bool ok = row < M;
            lo[i] = ok ? cur[row] : 0;
            ln[i] = ok ? nw[row] : 0;
        }
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            if (lo[i] != ln[i]) {
                const int r = threadIdx.x + i * APPLY_THREADS;       // r < 4096 < 2^16
                cur[row_base + r] = ln[i];
                s_list[atomicAdd(&s_n, 1)] = r | ((lo[i] & 0xff) << 16) | (ln[i] << 24);
                atomicAdd(&s_cin[ln[i]], 1);
                if (lo[i] >= 0) atomicAdd(&s_cout[lo[i]], 1);
            }
        }
    }
    __syncthreads();
    const int nchg = s_n;
    if (nchg == 0) continue;
    if (nchg <= APPLY_DIRECT_MAX) {
        // Few changes in this chunk (the late passes): the sorted path below would walk K clusters one after
        // the other with a dependent gather each; instead every (changed row, column) pair is handled by its
        // own thread -- independent loads, two global atomics (enter / leave).
        const float* xb = X + row_base * XP;
        for (int e = threadIdx.x; e < nchg * N; e += APPLY_THREADS) {
            const int i = e / N, j = e - i * N;
            const int v = s_list[i];
            const int o = (v >> 16) & 0xff, n = (v >> 24) & 0xff;
            const long long q = quant40(__ldg(xb + ((unsigned)(v & 0xffff) * (unsigned)XP + (unsigned)j)));
            atomicAdd(reinterpret_cast<unsigned long long*>(acc + (long long)(crow0 + n) * AP + j), (unsigned long long)q);
            if (o != 0xff)
                atomicAdd(reinterpret_cast<unsigned long long*>(acc + (long long)(crow0 + o) * AP + j), (unsigned long long)(-q));
        }
        if (threadIdx.x < K) {
            const int d = s_cin[threadIdx.x] - s_cout[threadIdx.x];
            if (d != 0)
                atomicAdd(reinterpret_cast<unsigned long long*>(acc + (long long)(crow0 + threadIdx.x) * AP + XP), (unsigned long long)(long long)d);
        }
        if (threadIdx.x == 0)
            atomicAdd(reinterpret_cast<unsigned long long*>(state + (long long)run * kStateWordsTc + LEIDA_KM_ST_CHANGED),
                      (unsigned long long)nchg);
        continue;
    }
    if (threadIdx.x == 0) {          // exclusive prefix sums (K <= 64)
        int a = 0, b = 0;
        for (int k = 0; k < K; ++k) {
            const int ci = s_cin[k], co = s_cout[k];
            s_pin[k] = a; s_pout[k] = b;
            s_cin[k] = a; s_cout[k] = b;
            a += ci; b += co;
        }
        s_cin[K] = a; s_cout[K] = b;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nchg; e += APPLY_THREADS) {
        const int v = s_list[e];
        const int o = (v >> 16) & 0xff, n = (v >> 24) & 0xff;
        s_in[atomicAdd(&s_pin[n], 1)] = (unsigned short)(v & 0xffff);
        if (o != 0xff) s_out[atomicAdd(&s_pout[o], 1)] = (unsigned short)(v & 0xffff);
    }
    __syncthreads();
    // thread (g, j4): columns 4 j4 .. 4 j4 + 3 of its share (rows g, g + groups, ...) of every cluster's lists
    const int C4 = (N + 3) >> 2;                                   // float4 columns (pad columns of X are zero)
    const int CW = C4 >= APPLY_THREADS ? APPLY_THREADS : ((C4 + 31) & ~31);
    const int groups = APPLY_THREADS / CW;
    const int g = threadIdx.x / CW, j0 = threadIdx.x - g * CW;
    if (g < groups) {
        const float* xb = X + row_base * XP;
        asm volatile("" : "+l"(xb));     // one 64-bit base register: element address = one IMAD.WIDE on a 32-bit index
        for (int k = 0; k < K; ++k) {
            const int ib = s_cin[k], ie = s_cin[k + 1], ob = s_cout[k], oe = s_cout[k + 1];
            if (ib == ie && ob == oe) continue;
            for (int j4 = j0; j4 < C4; j4 += CW) {
                long long sin4[4] = {0, 0, 0, 0}, sout4[4] = {0, 0, 0, 0};
                quant_sum4<UNIT>(xb, (unsigned)(4 * j4), (unsigned)XP, s_in, ib + g, ie, groups, sin4);
                quant_sum4<UNIT>(xb, (unsigned)(4 * j4), (unsigned)XP, s_out, ob + g, oe, groups, sout4);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const long long sum = sin4[c] - sout4[c];
                    if (sum != 0)
                        atomicAdd(reinterpret_cast<unsigned long long*>(acc + (long long)(crow0 + k) * AP + 4 * j4 + c),
                                  (unsigned long long)sum);
                }
            }
        }
    }
    if (threadIdx.x < K) {
        const int d = (s_cin[threadIdx.x + 1] - s_cin[threadIdx.x]) - (s_cout[threadIdx.x + 1] - s_cout[threadIdx.x]);
        if (d != 0)
            atomicAdd(reinterpret_cast<unsigned long long*>(acc + (long long)(crow0 + threadIdx.x) * AP + XP), (unsigned long long)(long long)d);
    }
    if (threadIdx.x == 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(state + (long long)run * kStateWordsTc + LEIDA_KM_ST_CHANGED),
                  (unsigned long long)nchg);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2D bf16 tensor (rows, NPK) row-major; box = 64 columns (128 B) x 128 rows, 128-byte swizzle.
static int make_map(CUtensorMap* map, const void* ptr, long long rows, int NPK, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return LEIDA_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)NPK, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)NPK * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return LEIDA_ERR_CUDA; }
    return LEIDA_OK;
}

}  // namespace leida

using namespace leida;

// columns of the bf16 operands: N features + 1 constant "bias" feature, rounded up to whole k-blocks
extern "C" int leida_kmeans_tc_tile_columns(void) { return BN; }

extern "C" int64_t leida_kmeans_tc_pitch(int N) { return ((int64_t)N + 1 + BK - 1) / BK * BK; }

extern "C" int leida_kmeans_tc_split_rows(const float* X, int64_t M, int N, int64_t x_pitch, void* X1, void* X2,
                                          leida_stream_t stream) {
    LEIDA_REQUIRE(X && X1 && X2, "null pointer");
    LEIDA_REQUIRE(M >= 1 && N >= 1 && x_pitch >= N, "sizes");
    const int NPK = (int)leida_kmeans_tc_pitch(N);
    const long long n = M * NPK;
    k_split_rows<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, M, N, x_pitch, (__nv_bfloat16*)X1,
                                                                                 (__nv_bfloat16*)X2, NPK);
    return check_launch("k_split_rows");
}

extern "C" int leida_kmeans_tc_plan(int n_runs, const int32_t* run_k, const int32_t* run_crow0, const int64_t* state,
                                    const float* cent, const double* cnorm, int N, int64_t x_pitch, int32_t* plan,
                                    int32_t* tile_first, int32_t* tile_count, int32_t* prun, int32_t* pcol,
                                    int32_t* run_pcol, void* C1, void* C2, int pack, leida_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LEIDA_REQUIRE(run_k && run_crow0 && state && cent && cnorm && plan && tile_first && tile_count && prun && pcol &&
                      run_pcol && C1 && C2, "null pointer");
    LEIDA_REQUIRE(n_runs >= 1 && N >= 1 && x_pitch >= N, "sizes");
    LEIDA_REQUIRE(n_runs <= PLAN_MAX_RUNS, "n_runs <= 4096 per batch");
    const int NPK = (int)leida_kmeans_tc_pitch(N);
    k_plan<<<1, 256, 0, st>>>(n_runs, run_k, (const long long*)state, plan, tile_first, tile_count, prun, pcol, run_pcol);
    if (pack)
        k_pack_centroids<<<n_runs, 256, 0, st>>>(run_k, run_crow0, run_pcol, cent, cnorm, N, (int)x_pitch,
                                                 (__nv_bfloat16*)C1, (__nv_bfloat16*)C2, NPK);
    return check_launch("leida_kmeans_tc_plan", pack ? 2 : 1);
}

// Error bound of a tensor-core score, in units of |x| (the centroids are unit vectors, so sum|x_j c_j| <= |x|).
//   dropped split terms x2*c2, x*rc, rx*c                        3 * 2^-18       (rigorous)
//   fp32 normalisation of the centroid                            2^-22           (rigorous)
//   index packing in the epilogue (6 mantissa bits)               2^-17           (rigorous)
//   fp32 accumulation inside the tensor core over 3*NPK/16 MMAs   one rounding of 2^-24 of a running sum <= |x| per
//   MMA.  The accumulation order and rounding of tcgen05.mma are undocumented, so this term is MEASURED, not derived:
//   leida_kmeans_tc_scores dumps the raw scores and tools/tc_score_error.py compares them with fp64 on 7.5e7 scores
//   per case (profiles/r02_tc_score_error.txt): the largest total error (split + normalisation + accumulation) was
//   3.1e-6 / 2.9e-6 / 4.6e-6 |x| at N = 116 / 400 / 1000 -- it grows with the number of MMAs at about 0.26 * 2^-24 per
//   MMA on clustered rows (coherent partial sums; iid rows stay below 3e-6) -- so the model term is ~4x the observed
//   accumulation error, on top of the rigorous terms, which alone exceed every observed total.  The margin test
//   compares two scores (x 2): tau = 4.2e-5 / 4.9e-5 / 6.2e-5 at N = 116 / 400 / 1000 (round 1's worst-case model:
//   1.2e-4 / 3.0e-4 / 6.2e-4, which queued 10x more pairs at N = 400).  tests/test_gpu_scale.py asserts
//   4 * (observed max) <= tau and checks labels against fp64 pass by pass.
// On top of the tests the bound is watched in production: every re-checked pair carries its tensor-core argmax and
// margin, and LEIDA_KM_ST_TAU_WATCH records the largest margin (as a fraction of the threshold) at which the fp64
// argmin disagreed (observed: 0.1); the host warns when that fraction exceeds 1/2.
static float default_tau(int NPK) {
    const float rigorous = 3.0f * 3.8147e-6f + 2.4e-7f + 7.63e-6f;
    const float acc_term = 3.0f * (float)(NPK / 16) * 5.9605e-8f;
    return 2.0f * (rigorous + acc_term);
}

extern "C" double leida_kmeans_tc_tau(int N) { return (double)default_tau((int)leida_kmeans_tc_pitch(N)); }

namespace leida {
// Tensor maps depend only on (device, base pointer, rows, pitch): the per-pass launches reuse them.
struct MapKey { int dev; const void* ptr; long long rows; int npk; int box; };
struct MapSlot { MapKey k; CUtensorMap m; bool used; };
static MapSlot g_maps[32];
static int g_map_next = 0;
static std::mutex g_map_mu;
static bool g_attr_set[64][2];

static int cached_map(CUtensorMap* out, int dev, const void* ptr, long long rows, int npk, int box) {
    std::lock_guard<std::mutex> lk(g_map_mu);
    for (auto& s : g_maps)
        if (s.used && s.k.dev == dev && s.k.ptr == ptr && s.k.rows == rows && s.k.npk == npk && s.k.box == box) { *out = s.m; return LEIDA_OK; }
    MapSlot& s = g_maps[g_map_next];
    g_map_next = (g_map_next + 1) % 32;
    const int rc = make_map(&s.m, ptr, rows, npk, box);
    if (rc) { s.used = false; return rc; }
    s.k = MapKey{dev, ptr, rows, npk, box};
    s.used = true;
    *out = s.m;
    return LEIDA_OK;
}

static int launch_tc_assign(const void* X1, const void* X2, const double* xnorm, int64_t M, int N, const void* C1,
                            const void* C2, int64_t c_rows, const int32_t* plan, const int32_t* tile_first,
                            const int32_t* tile_count, const int32_t* prun, int32_t* labels_new, int32_t* worklist,
                            int64_t worklist_cap, double tau, float* dump, int64_t dump_pitch, cudaStream_t st) {
    LEIDA_REQUIRE(X1 && X2 && xnorm && C1 && C2 && plan && tile_first && tile_count && prun && labels_new && worklist,
                  "null pointer");
    LEIDA_REQUIRE(M >= 1 && M < (1LL << 31) && N >= 1 && c_rows >= BN, "sizes");
    const int NPK = (int)leida_kmeans_tc_pitch(N);
    int dev = 0;
    LEIDA_CUDA_TRY(cudaGetDevice(&dev));
    CUtensorMap mx1, mx2, mc1, mc2;
    int rc;
    if ((rc = cached_map(&mx1, dev, X1, M, NPK, BM)) || (rc = cached_map(&mx2, dev, X2, M, NPK, BM)) ||
        (rc = cached_map(&mc1, dev, C1, c_rows, NPK, BN)) || (rc = cached_map(&mc2, dev, C2, c_rows, NPK, BN)))
        return rc;
    TcParams p;
    p.M = M; p.nkb = NPK / BK; p.plan = plan; p.tile_first = tile_first; p.tile_count = tile_count;
    p.prun = prun; p.xnorm = xnorm;
    p.tau = tau > 0.0 ? (float)tau : default_tau(NPK);
    p.labels_new = labels_new; p.worklist = worklist; p.worklist_cap = worklist_cap;
    p.dump = dump; p.dump_pitch = dump_pitch;
    const size_t smem = 1024 + (size_t)ST * STAGE_BYTES + 32 * sizeof(uint64_t);
    const int which = dump ? 1 : 0;
    if (dev < 64 && !g_attr_set[dev][which]) {      // once per device; keeps the call legal inside stream capture
        if (dump) LEIDA_CUDA_TRY(cudaFuncSetAttribute(k_tc_assign<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else LEIDA_CUDA_TRY(cudaFuncSetAttribute(k_tc_assign<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g_attr_set[dev][which] = true;
    }
    const int n_rt = (int)((M + BM - 1) / BM);
    const int grid = n_rt < sm_count() ? n_rt : sm_count();
    if (dump) k_tc_assign<true><<<grid, TC_THREADS, smem, st>>>(mx1, mx2, mc1, mc2, p);
    else k_tc_assign<false><<<grid, TC_THREADS, smem, st>>>(mx1, mx2, mc1, mc2, p);
    return check_launch("k_tc_assign");
}
}  // namespace leida

extern "C" int leida_kmeans_tc_assign(const void* X1, const void* X2, const double* xnorm, int64_t M, int N, const void* C1,
                                      const void* C2, int64_t c_rows, const int32_t* plan, const int32_t* tile_first,
                                      const int32_t* tile_count, const int32_t* prun, int32_t* labels_new,
                                      int32_t* worklist, int64_t worklist_cap, double tau, leida_stream_t stream) {
    return launch_tc_assign(X1, X2, xnorm, M, N, C1, C2, c_rows, plan, tile_first, tile_count, prun, labels_new, worklist,
                            worklist_cap, tau, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int leida_kmeans_tc_scores(const void* X1, const void* X2, const double* xnorm, int64_t M, int N, const void* C1,
                                      const void* C2, int64_t c_rows, const int32_t* plan, const int32_t* tile_first,
                                      const int32_t* tile_count, const int32_t* prun, int32_t* labels_new,
                                      int32_t* worklist, int64_t worklist_cap, float* scores, int64_t score_pitch,
                                      leida_stream_t stream) {
    LEIDA_REQUIRE(scores && score_pitch >= c_rows, "scores (M, score_pitch >= c_rows)");
    return launch_tc_assign(X1, X2, xnorm, M, N, C1, C2, c_rows, plan, tile_first, tile_count, prun, labels_new, worklist,
                            worklist_cap, 0.0, scores, score_pitch, (cudaStream_t)stream);
}

extern "C" int leida_kmeans_tc_apply(const float* X, int64_t M, int N, int64_t x_pitch, int unit_rows, int n_runs_bound,
                                     const int32_t* plan, const int32_t* prun, const int32_t* run_k,
                                     const int32_t* run_crow0, int32_t* labels_cur, const int32_t* labels_new,
                                     int64_t* acc, int64_t* state, leida_stream_t stream) {
    LEIDA_REQUIRE(X && plan && prun && run_k && run_crow0 && labels_cur && labels_new && acc && state, "null pointer");
    LEIDA_REQUIRE(M >= 1 && N >= 1 && x_pitch >= N && x_pitch < (1 << 19) && n_runs_bound >= 1 && n_runs_bound <= 65535, "sizes");
    dim3 grid((unsigned)(n_runs_bound < 48 ? n_runs_bound : 48), (unsigned)((M + APPLY_ROWS - 1) / APPLY_ROWS));
    if (unit_rows)
        k_diff_apply<true><<<grid, APPLY_THREADS, 0, (cudaStream_t)stream>>>(X, M, N, (int)x_pitch, plan, prun, run_k, run_crow0,
                                                                             labels_cur, labels_new, (long long*)acc,
                                                                             (long long*)state);
    else
        k_diff_apply<false><<<grid, APPLY_THREADS, 0, (cudaStream_t)stream>>>(X, M, N, (int)x_pitch, plan, prun, run_k, run_crow0,
                                                                              labels_cur, labels_new, (long long*)acc,
                                                                              (long long*)state);
    return check_launch("k_diff_apply");
}
