#!/usr/bin/env python
"""bench.py -- end-to-end LEiDA hot path throughput (volumes/s) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] # the reference algorithm on the host CPUs

One "step" = one pass of the whole hot path over one batch of synthetic input:
    Hilbert phase -> leading eigenvector -> cosine k-means for K=2..20 x 15 restarts -> dynamics metrics.
Workload at N=1: BASELINE.json configs[1] (100 subjects x 116 ROIs x 1200 volumes).  With N>1 every
rank holds one such shard (weak scaling: 100*N subjects; k-means is ONE clustering over all N*119800
rows with the int64 centroid sums all-reduced over NCCL every Lloyd iteration).

`value` : volumes/s with the BOLD series already resident in HBM (CUDA events, max over ranks).
`e2e`   : same metric through the public API (Leida.from_arrays(...).fit_predict) from HOST numpy
          arrays: pinned H2D copies of the inputs and D2H of eigenvectors, labels, centroids and
          metric tables are inside the timed region.
`roofline`: the dominant kernel (k_assign) timed live with CUDA events around every launch inside the
          timed region; achieved = algorithmic bytes ((4*M*N + 8*M) per run per pass) / time.
`cpu_baseline` / `--impl reference`: the oracle port of the reference algorithm on the host cores, on a
          bounded sample of the same workload (stated in `sample`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))

import numpy as np  # noqa: E402

WORKLOAD = dict(name="config2: 100 subjects x 116 ROIs x 1200 volumes per GPU, K=2..20, n_init=15",
                S=100, N=116, T=1200, K_min=2, K_max=20, n_init=15, kind="states", TR=0.72)


def load_synthetic():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "leida_synthetic", os.path.join(ROOT, "leida-python_b200", "pyleida", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak():
    """Dense bf16 TFLOP/s: the burst figure, because k_tc_assign launches are short and isolated."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if "bf16_tflops" in d:
            return float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, cuBLAS burst)"
    return 2250.0, "nominal dense bf16 (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
_G = {}


def _fit_one(task):
    k, idx = task
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import leida_oracle as O
    km = O.KMeansLeida(k=k, n_init=1, n_iter=1000)
    km.fit(_G["X"], init_indices=[idx])
    return k, float(km.inertia_), km.labels_, km.n_iter_runs_[0]


def _phase_one(x):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import leida_oracle as O
    return O.leading_eigvec_rank2(O.hilbert_phase(x))


def run_cpu_sample(wl, sample_subjects, n_init_sample, steps, warmup, cores):
    syn = load_synthetic()
    ts, classes = syn.make_shard(wl["S"], wl["N"], wl["T"], 0, sample_subjects, wl["kind"])
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        r = _cpu_step_with_pool(ts, classes, wl, n_init_sample, cores)
        r["wall"] = time.perf_counter() - t0
        if it >= warmup:
            times.append(r)
    return times


def _cpu_step_with_pool(ts, classes, wl, n_init_sample, cores):
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import leida_oracle as O
    syn = load_synthetic()
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    if cores > 1:
        with ctx.Pool(cores) as pool:
            eig = pool.map(_phase_one, [np.asarray(v) for v in ts.values()])
    else:
        eig = [_phase_one(np.asarray(v)) for v in ts.values()]
    X = np.array(np.vstack(eig), dtype=np.float32)
    t1 = time.perf_counter()
    inits = syn.init_indices(X.shape[0], wl["K_min"], wl["K_max"], n_init_sample, seed=0)
    tasks = [(k, idx) for k in range(wl["K_min"], wl["K_max"] + 1) for idx in inits[k]]
    tasks.sort(key=lambda t: -t[0])
    _G["X"] = X                                   # inherited by the forked workers
    if cores > 1:
        with ctx.Pool(cores) as pool:
            results = pool.map(_fit_one, tasks, chunksize=1)
    else:
        results = [_fit_one(t) for t in tasks]
    best = {}
    for k, inertia, labels, _ in results:
        if k not in best or inertia < best[k][0]:
            best[k] = (inertia, labels)
    t2 = time.perf_counter()
    nv = [v.shape[1] - 2 for v in ts.values()]
    subs = np.repeat(list(ts.keys()), nv)
    conds = np.repeat([classes[s][0] for s in ts], nv)
    for k in best:
        km = O.KMeansLeida(k=k)
        km.labels_, km.cluster_centers_ = best[k][1], np.zeros((k, X.shape[1]), dtype=np.float32)
        km._remap_labels()
        O.dynamics_tables(subs, conds, km.labels_, TR=wl["TR"])
    t3 = time.perf_counter()
    return dict(stage12=t1 - t0, kmeans=t2 - t1, metrics=t3 - t2, M=int(X.shape[0]),
                lloyd_iters=int(sum(r[3] for r in results)))


def cpu_value(times, wl, n_init_sample):
    """volumes/s of the FULL per-volume work: the k-means time of the sample is scaled by
    n_init / n_init_sample (restarts are independent and equally expensive on average)."""
    scale = wl["n_init"] / float(n_init_sample)
    t = np.mean([r["stage12"] + r["kmeans"] * scale + r["metrics"] for r in times])
    return times[0]["M"] / t, t


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOAD
    cores = os.cpu_count() or 1
    sample_subjects, n_init_sample = 16, 2
    times = run_cpu_sample(wl, sample_subjects, n_init_sample, args.steps, args.warmup, cores)
    value, t_equiv = cpu_value(times, wl, n_init_sample)
    sample = (f"oracle port (oracle/leida_oracle.py: numpy/scipy restatement of the reference algorithm, closed-form "
              f"eigenvector), {sample_subjects} of {wl['S']} subjects x {wl['N']} ROIs x {wl['T']} volumes, K=2..20, "
              f"{n_init_sample} of {wl['n_init']} restarts per K with the k-means time scaled x{wl['n_init'] / n_init_sample:g}; "
              f"(K, restart) runs spread over {cores} processes")
    out = {
        "impl": "reference", "metric": "volumes/sec end-to-end (phase->eigvec->k-means K=2..20->metrics)",
        "value": value, "unit": "volumes/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean([r["wall"] for r in times])), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 phases/eigenvectors/distances, f32 rows/centroids (as the reference)",
        "data": "synthetic", "config": {"workload": wl["name"], "kind": wl["kind"]},
        "cpu_baseline": {"value": value, "unit": "volumes/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "stage_seconds_sample": {k: float(np.mean([r[k] for r in times])) for k in ("stage12", "kmeans", "metrics")},
    }
    emit(out)
    return 0


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def native_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import pyleida
    from pyleida import _native as nvl
    from pyleida.dynamics_metrics.metrics import Segments
    from pyleida.parallel import Comm
    from pyleida.pipeline import run_device
    syn = load_synthetic()
    wl = WORKLOAD
    comm = Comm()
    S_total = wl["S"] * world
    lo, hi = rank * wl["S"], (rank + 1) * wl["S"]
    ts, classes = syn.make_shard(S_total, wl["N"], wl["T"], lo, hi, wl["kind"])
    ids = list(ts.keys())
    Tp = wl["T"] - 2
    M_local, M_global = wl["S"] * Tp, S_total * Tp
    inits = syn.init_indices(M_global, wl["K_min"], wl["K_max"], wl["n_init"], seed=0)
    seg = Segments.contiguous(np.asarray(ids), np.asarray([classes[s][0] for s in ids]), [Tp] * len(ids))
    x_host = np.stack([ts[s] for s in ids])
    x_dev = torch.from_numpy(x_host).to(dev)
    seg_off = torch.from_numpy(seg.off).to(dev)
    seg_rows = torch.from_numpy(seg.rows).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(profile=None):
        flush.zero_()                                  # L2 flush between steps (256 MB > 126 MB L2)
        return run_device([x_dev], wl["N"], seg_off, seg_rows, K_min=wl["K_min"], K_max=wl["K_max"],
                          n_init=wl["n_init"], init_indices=inits, keep_f64=True, comm=comm, profile=profile)

    def step_e2e():
        flush.zero_()
        m = pyleida.Leida.from_arrays(ts, classes, comm=comm)
        m.fit_predict(TR=wl["TR"], K_min=wl["K_min"], K_max=wl["K_max"], n_init=wl["n_init"], init_indices=inits,
                      keep_eigenvectors=True, scores=False, stats=False)     # the metric covers phase->eigvec->k-means->metrics
        # consume the result like a user would: labels of every K are already host arrays; the metric
        # DataFrames are assembled on access -- touch the three tables of one K
        k = wl["K_max"]
        _ = (m.predictions[f"k_{k}"].values[-1], m.occupancies(k).shape, m.dwell_times(k).shape, m.transitions(k).shape)
        return m

    for _ in range(args.warmup):
        step_device()                                  # result dropped at once: the timed steps reuse its memory
    profile = {"*": True} if os.environ.get("LEIDA_BENCH_PROFILE_ALL") else {"assign": []}
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = nvl.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph_launches = 0
    step_marks = []
    stats = None
    for _ in range(args.steps):
        res = step_device(profile)
        graph_launches += res.batch.graph_launches
        b = res.batch
        stats = dict(run_passes=b.total_passes(), lockstep=b.lockstep_passes, n_runs=len(b._ks),
                     sum_k_passes=float(sum(b.k(i) * b.passes(i) for i in range(len(b._ks)))),
                     rechecked=b.rechecked_total, changed=b.changed_total)
        del res, b                                     # free the step's tensors before the next step allocates
        m = torch.cuda.Event(enable_timing=True)
        m.record()
        step_marks.append(m)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = nvl.launch_count() - l0 + graph_launches     # kernels inside replayed CUDA graphs are counted by the engine
    ms = e0.elapsed_time(e1)
    per_step_ms = [a.elapsed_time(b) for a, b in zip([e0] + step_marks[:-1], step_marks)]
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    ln = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    value = M_global * args.steps / (ms * 1e-3)

    # roofline of the dominant kernel, from the event pairs recorded inside the timed region
    kernel_ms = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in profile.items() if k != "*"}
    assign_ms = kernel_ms.get("assign", 0.0)
    n_assign = len(profile.get("assign", []))
    run_passes = stats["run_passes"]                    # per step (same every step: deterministic)
    alg_bytes_per_run_pass = 4.0 * M_local * wl["N"] + 8.0 * M_local
    alg_bytes = alg_bytes_per_run_pass * run_passes * args.steps
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (assign_ms * 1e-3) / 1e9 if assign_ms > 0 else 0.0
    # what the tensor-core kernel really does per launch: one read of the bf16 x 2 split rows for ALL
    # active runs, and 3 bf16 MMAs (X1*C1 + X1*C2 + X2*C1) per output
    npk = int(nvl.lib.leida_kmeans_tc_pitch(wl["N"]))
    sum_k_passes = stats["sum_k_passes"]
    tensor_flops = 3 * 2.0 * M_local * npk * sum_k_passes * args.steps
    actual_bytes = n_assign * (2 * 2.0 * M_local * npk) + 4.0 * M_local * run_passes * args.steps
    roofline = {"kernel": "k_tc_assign (tcgen05 bf16x3 distance GEMM + top-2 epilogue)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch that serves all 285 runs of this
                # workload (ncu --set full, profiles/r01_k_tc_assign_ncu_full.txt capture E: 64.6 MB read =
                # one pass over the bf16x2 rows, 89.2 MB written = the int32 labels); that launch's algorithmic
                # bytes on the SURVEY metric are 285 x 56.5 MB = 16.1 GB
                "traffic": 153.8e6 if (world == 1 and wl["S"] == 100 and wl["N"] == 116) else None,
                "traffic_launch": "full pass, 285 active runs (captured once with ncu; per-launch traffic shrinks with the active-run count)",
                "peak_source": peak_src, "launches": n_assign, "avg_launch_ms": assign_ms / max(n_assign, 1),
                "share_of_step": assign_ms / ms,
                "algorithmic_bytes_per_launch": alg_bytes / max(n_assign, 1),
                "actual_bytes_per_launch_model": actual_bytes / max(n_assign, 1),
                "tensor_tflops_bf16": tensor_flops / (assign_ms * 1e-3) / 1e12 if assign_ms > 0 else 0.0,
                "tensor_peak_tflops_bf16": tensor_peak()[0], "tensor_peak_source": tensor_peak()[1],
                "tensor_frac": (tensor_flops / (assign_ms * 1e-3) / 1e12 / tensor_peak()[0]) if assign_ms > 0 else 0.0,
                "other_kernels_ms_per_step": {k: v / args.steps for k, v in kernel_ms.items() if k != "assign"},
                "note": "algorithmic bytes = (4*M*N + 8*M) per run per Lloyd pass (SURVEY 8d). One launch serves every "
                        "active run of the pass from ONE read of the rows (a tensor-core contraction over all runs' "
                        "centroids), so achieved exceeds the HBM peak by design; actual_bytes_per_launch_model is what "
                        "a launch really moves (bf16x2 rows once + int32 labels per active run)"}

    # end-to-end through the public API from host arrays
    for _ in range(2):
        step_e2e()                                     # warm-up (pinned staging buffers, allocator); result dropped
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e2e = max(1, min(args.steps, 5))
    f0.record()
    t_wall = time.perf_counter()
    h2d = d2h = 0
    for _ in range(n_e2e):
        m = step_e2e()
        h2d, d2h = int(m.h2d_bytes), int(m.d2h_bytes)
        del m
    f1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    e2e_ms = max(f0.elapsed_time(f1), 1e3 * t_wall)
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = M_global * n_e2e / (float(te.item()) * 1e-3)

    out = None
    if rank == 0:
        out = {
            "metric": "volumes/sec end-to-end (phase->eigvec->k-means K=2..20->metrics)",
            "value": value, "unit": "volumes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 phases/eigenvectors/distance re-check, f32 rows/centroids/dot products, int64 centroid sums",
            "data": "synthetic",
            "config": {"workload": wl["name"], "kind": wl["kind"], "subjects_total": S_total, "volumes_total": M_global,
                       "kmeans_runs": len(inits) * wl["n_init"], "kmeans_run_passes_per_step": run_passes,
                       "lockstep_passes_per_step": stats["lockstep"],
                       "fp64_rechecks_per_step": stats["rechecked"], "label_changes_per_step": stats["changed"],
                       "init": "explicit init rows from default_rng(0), shared with the CPU arm",
                       "e2e_call": "Leida.from_arrays(ts, classes).fit_predict(..., scores=False, stats=False): the metric is "
                                   "phase->eigvec->k-means->metrics (SURVEY 8d); quality scores and permutation tests are separate stages",
                       "l2_flush_between_steps": True, "per_step_ms": [round(v, 2) for v in per_step_ms], "parallelism": f"subjects sharded over {world} GPU(s)"},
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": "volumes/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "steps": n_e2e,
                    "ms_per_step": float(te.item()) / n_e2e},
            "gpu_launches": int(ln.item()),
            "clocks": clocks,
        }
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample_subjects, n_init_sample = 16, 2
        times = run_cpu_sample(wl, sample_subjects, n_init_sample, 1, 0, cores)
        cv, _ = cpu_value(times, wl, n_init_sample)
        out["cpu_baseline"] = {
            "value": cv, "unit": "volumes/s", "cores": cores, "kind": "port",
            "sample": (f"oracle port of the reference algorithm, {sample_subjects} of {wl['S']} subjects at full N,T, K=2..20, "
                       f"{n_init_sample} of {wl['n_init']} restarts per K (k-means time scaled x{wl['n_init'] / n_init_sample:g}), "
                       f"runs spread over {cores} processes; {times[0]['wall']:.1f} s of CPU wall time")}
    if rank == 0:
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_RESULT_FD = None


def emit(obj):
    """The ONE JSON line of the contract, on the process's original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


def main():
    # Libraries print to stdout on their own (NCCL's version banner at communicator creation): route
    # everything that is not the result line to stderr.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return native_arm(args)


if __name__ == "__main__":
    sys.exit(main())
