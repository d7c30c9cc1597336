#!/usr/bin/env python
"""bench.py -- end-to-end LEiDA hot path throughput (volumes/s) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config config3]   # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]      # the reference algorithm on the host CPUs

One "step" = one pass of the whole hot path over one batch of synthetic input:
    Hilbert phase -> leading eigenvector -> cosine k-means for every K of the sweep x n_init restarts -> dynamics metrics.
Workload (default): BASELINE.json configs[2], the configuration the metric is quoted on --
    1000 subjects x 400 ROIs x 1200 volumes, K=2..20, 15 restarts -- as ONE fixed data set whose subjects are
    sharded over the --gpus N ranks (strong scaling).  k-means is one clustering over all 1 198 000 rows; the int64
    centroid sums are exchanged between the ranks every Lloyd iteration.
    --config config2 | config4 | config5 | config1 select the other BASELINE configurations; at N=1 the default run also
    carries a short config-2 measurement as `secondary`.

`value`   : volumes/s with the BOLD series already resident in HBM (CUDA events, max over ranks).
`e2e`     : same metric through the public API (Leida.from_arrays(...).fit_predict) from HOST numpy arrays: pinned H2D
            copies of the inputs and D2H of eigenvectors, labels, centroids and metric tables inside the timed region.
`roofline`: the dominant kernel (k_tc_assign), timed live with CUDA events around every launch inside the timed
            region.  bound = "tensor": achieved = issued bf16 tensor-core FLOP/s (3 MMAs per output: bf16 x 3 split),
            peak = measured cuBLAS bf16 (MEASURED_PEAKS.json).  The SURVEY 8d per-run HBM figure is kept under
            `survey_metric`; `kernels` gives every other kernel of the pass loop its own time and achieved bytes.
`parity`  : computed outside the timed region on the last step's results: for several (K, restart) runs, one assignment
            of ALL rows against the run's final centroids with an independent fp64 argmin (torch fp64 matmul, the
            reference's formula 1 - x.c/(|x||c|), first minimum) -- must equal the run's final labels (Lloyd fixed
            point); plus a position-weighted hash of the final labels of every K, identical for any --gpus N.
`cpu_baseline` / `--impl reference`: the oracle port of the reference algorithm on the host cores, on a bounded sample of
            the same workload (stated in `sample`), EXTRAPOLATED to the full restart count.
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

METRIC = "volumes/sec end-to-end (phase->eigvec->k-means K sweep->metrics)"
WORKLOADS = {
    "config1": dict(name="config1: 20 subjects x 90 ROIs (AAL) x 200 volumes, K=2..20, n_init=15",
                    S=20, N=90, T=200, K_min=2, K_max=20, n_init=15, kind="states", TR=0.72),
    "config2": dict(name="config2: 100 subjects x 116 ROIs x 1200 volumes (HCP-like), K=2..20, n_init=15",
                    S=100, N=116, T=1200, K_min=2, K_max=20, n_init=15, kind="states", TR=0.72),
    "config3": dict(name="config3: 1000 subjects x 400 ROIs (Schaefer-400) x 1200 volumes, K=2..20, n_init=15, "
                         "one fixed data set sharded over the GPUs",
                    S=1000, N=400, T=1200, K_min=2, K_max=20, n_init=15, kind="states", TR=0.72),
    "config4": dict(name="config4: 200 subjects x 1000 ROIs x 1200 volumes (general eigh path stress), K=2..20, n_init=15",
                    S=200, N=1000, T=1200, K_min=2, K_max=20, n_init=15, kind="states", TR=0.72),
    "config5": dict(name="config5: 1000 subjects x 400 ROIs x 1200 volumes, K=2..40, n_init=50",
                    S=1000, N=400, T=1200, K_min=2, K_max=40, n_init=50, kind="states", TR=0.72),
}
DEFAULT_CONFIG = "config3"
DTYPE = "f64 phases/eigenvectors/distance re-check, f32 rows/centroids, bf16x3 tensor-core dot products (fp32-grade), int64 centroid sums"


def static_config(key, wl):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": wl["name"], "config": key, "kind": wl["kind"], "subjects": wl["S"], "rois": wl["N"],
            "volumes_per_subject": wl["T"], "rows": wl["S"] * (wl["T"] - 2), "K_min": wl["K_min"], "K_max": wl["K_max"],
            "n_init": wl["n_init"], "kmeans_runs": (wl["K_max"] - wl["K_min"] + 1) * wl["n_init"],
            "init": "explicit init rows from default_rng(0), shared by both arms",
            "l2_flush_between_steps": True, "sharding": "subjects over ranks (strong scaling: the data set is fixed)"}


def load_synthetic():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "leida_synthetic", os.path.join(ROOT, "leida-python_b200", "pyleida", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def hbm_peak():
    d = _peaks()
    if "hbm_gbs" in d:
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak(long_step=False):
    """Dense bf16 TFLOP/s measured on this pool (MEASURED_PEAKS.json): the burst figure for kernels timed alone or inside
    short steps, the sustained one for a kernel timed inside a long, power-capped step (the config-3 step runs the
    tensor cores for ~0.6 s of every second and nvidia-smi reports sw_power_cap throughout)."""
    d = _peaks()
    if long_step and "bf16_tflops_sustained" in d:
        return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained, cuBLAS back to back)"
    if "bf16_tflops" in d:
        return float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, cuBLAS burst)"
    return 2250.0, "nominal dense bf16 (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
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
CPU_SAMPLE = {  # config -> (subjects, restarts per K) of the bounded CPU sample
    "config1": (20, 2), "config2": (16, 2), "config3": (12, 1), "config4": (6, 1), "config5": (12, 1),
}


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
    return dict(stage12=t1 - t0, kmeans=t2 - t1, metrics=t3 - t2, M=int(X.shape[0]), runs=len(results),
                lloyd_iters=int(sum(r[3] for r in results)))


def cpu_value(times, wl, n_init_sample, full_passes_per_run=None):
    """volumes/s of the FULL per-volume work, EXTRAPOLATED from the sample: the k-means time of the sample is scaled by
    n_init / n_init_sample (restarts are independent and equally expensive on average).  When the full-size run's
    passes per run are known (the GPU arm of the same process), a second figure also scales the k-means time by
    (passes per run at full size) / (Lloyd iterations per run of the sample): Lloyd needs more iterations on more rows."""
    scale = wl["n_init"] / float(n_init_sample)
    t = np.mean([r["stage12"] + r["kmeans"] * scale + r["metrics"] for r in times])
    v = times[0]["M"] / t
    v_iter = None
    if full_passes_per_run:
        it_sample = np.mean([r["lloyd_iters"] / r["runs"] for r in times]) + 1.0      # + the initial assignment
        t2 = np.mean([r["stage12"] + r["kmeans"] * scale * full_passes_per_run / it_sample + r["metrics"] for r in times])
        v_iter = times[0]["M"] / t2
    return v, t, v_iter


def cpu_sample_text(key, wl, sample_subjects, n_init_sample, cores, wall=None):
    s = (f"EXTRAPOLATED from a bounded sample: oracle port (oracle/leida_oracle.py: numpy/scipy restatement of the reference "
         f"algorithm, closed-form eigenvector, vectorised -- a much faster baseline than the as-is reference), "
         f"{sample_subjects} of {wl['S']} subjects x {wl['N']} ROIs x {wl['T']} volumes, K={wl['K_min']}..{wl['K_max']}, "
         f"{n_init_sample} of {wl['n_init']} restarts per K with the k-means time scaled x{wl['n_init'] / n_init_sample:g}; "
         f"(K, restart) runs spread over {cores} processes; volumes/s = sample rows / extrapolated time")
    if wall is not None:
        s += f"; {wall:.1f} s of CPU wall time per sample step"
    return s


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    key, wl = args.config, WORKLOADS[args.config]
    cores = os.cpu_count() or 1
    sample_subjects, n_init_sample = CPU_SAMPLE[key]
    times = run_cpu_sample(wl, sample_subjects, n_init_sample, args.steps, args.warmup, cores)
    value, t_equiv, _ = cpu_value(times, wl, n_init_sample)
    sample = cpu_sample_text(key, wl, sample_subjects, n_init_sample, cores, float(np.mean([r["wall"] for r in times])))
    out = {
        "impl": "reference", "metric": METRIC,
        "value": value, "unit": "volumes/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean([r["wall"] for r in times])), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64 phases/eigenvectors/distances, f32 rows/centroids (as the reference)",
        "data": "synthetic", "config": static_config(key, wl),
        "cpu_baseline": {"value": value, "unit": "volumes/s", "cores": cores, "kind": "port", "sample": sample,
                         "extrapolated": True},
        "e2e": {"value": value, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "stage_seconds_sample": {k: float(np.mean([r[k] for r in times])) for k in ("stage12", "kmeans", "metrics")},
        "lloyd_iterations_per_run_sample": float(np.mean([r["lloyd_iters"] / r["runs"] for r in times])),
    }
    emit(out)
    return 0


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def make_inputs(syn, wl, lo, hi, threads=8):
    """Subjects lo..hi-1 of the workload (seeded per subject), generated on a few threads."""
    from concurrent.futures import ThreadPoolExecutor
    cuts = np.linspace(lo, hi, min(threads, max(1, hi - lo)) + 1).astype(int)
    with ThreadPoolExecutor(max_workers=threads) as pool:
        parts = list(pool.map(lambda ab: syn.make_shard(wl["S"], wl["N"], wl["T"], int(ab[0]), int(ab[1]), wl["kind"]),
                              zip(cuts[:-1], cuts[1:])))
    ts, classes = {}, {}
    for t, c in parts:
        ts.update(t)
        classes.update(c)
    return ts, classes


def fp64_parity(torch, res, wl, comm, row_lo, picks):
    """Independent fp64 check of the Lloyd fixed point (see the module docstring).  Returns a dict."""
    eng, b = res.engine, res.batch
    N = eng.N
    M = eng.M
    mism = ties = 0
    checked = []
    for i in picks:
        k = b.k(i)
        C = b.centers_device(i).double()                        # (k, N) fp32 centroids of the run, as fp64
        cn = C.norm(dim=1)
        lab = b.labels_device(i)
        bad = tie = 0
        for r0 in range(0, M, 1 << 17):
            X = eng.X[r0:r0 + (1 << 17), :N].double()
            xn = X.norm(dim=1)
            d = 1.0 - (X @ C.T) / (xn[:, None] * cn[None, :])
            ref = d.argmin(dim=1)                                # first minimum
            got = lab[r0:r0 + (1 << 17)].long()
            ne = ref != got
            if bool(ne.any()):
                idx = ne.nonzero().squeeze(1)
                gap = (d[idx, got[idx]] - d[idx, ref[idx]]).abs()
                tie += int((gap <= 1e-13).sum())                 # fp64 rounding-level ties: either label is "the" argmin
                bad += int((gap > 1e-13).sum())
        mism += bad
        ties += tie
        checked.append({"k": k, "rows": M, "passes": b.passes(i), "mismatches": bad, "fp64_ties": tie})
    # position-weighted hash of the final (relabelled) labels of every K: identical for any number of ranks
    lab_all = res.labels.long()                                  # (n_K, M_local)
    g = torch.arange(M, device=lab_all.device, dtype=torch.int64) + int(row_lo)
    h = 0
    for c in range(lab_all.shape[0]):
        w = (g * 31 + (c + 2) * 17) % 1000003
        h += int(((lab_all[c] + 1) * w).sum().item())
    t = torch.tensor([mism, ties, h], dtype=torch.int64, device=lab_all.device)
    if comm.world > 1:
        comm.all_reduce_sum_(t)
    inert = {f"k_{k}": repr(float(res.models[k].inertia_)) for k in (wl["K_min"], wl["K_max"])}
    return {"check": "fp64 argmin (torch fp64 matmul) of every row against the run's final centroids == final labels",
            "runs_checked": checked if comm.world == 1 else len(checked), "mismatches": int(t[0].item()),
            "fp64_ties": int(t[1].item()), "labels_hash": int(t[2].item()), "inertia": inert,
            "tau": b.tau, "tau_watch": b.tau_watch}


def measure(torch, dist, args, key, wl, steps, warmup, with_e2e=True, with_parity=True):
    """Device-resident + end-to-end measurement of one workload on the current process group."""
    import pyleida
    from pyleida import _native as nvl
    from pyleida.dynamics_metrics.metrics import Segments
    from pyleida.parallel import Comm, shard_bounds
    from pyleida.pipeline import run_device
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    syn = load_synthetic()
    comm = Comm()
    lo, hi = shard_bounds(wl["S"], world, rank)
    t_gen = time.perf_counter()
    ts, classes = make_inputs(syn, wl, lo, hi)
    t_gen = time.perf_counter() - t_gen
    ids = list(ts.keys())
    Tp = wl["T"] - 2
    M_local, M_global = (hi - lo) * Tp, wl["S"] * Tp
    row_lo = lo * Tp
    inits = syn.init_indices(M_global, wl["K_min"], wl["K_max"], wl["n_init"], seed=0)
    seg = Segments.contiguous(np.asarray(ids), np.asarray([classes[s][0] for s in ids]), [Tp] * len(ids))
    x_dev = torch.empty((len(ids), wl["N"], wl["T"]), dtype=torch.float64, device=dev)
    for i, s in enumerate(ids):
        x_dev[i].copy_(torch.from_numpy(ts[s]))
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
        # consume the result like a user would: labels of every K are host arrays; the metric DataFrames are assembled on
        # access -- build the three tables of every K (the CPU arm builds them all too)
        for k in range(wl["K_min"], wl["K_max"] + 1):
            _ = (m.predictions[f"k_{k}"].values[-1], m.occupancies(k).shape, m.dwell_times(k).shape, m.transitions(k).shape)
        return m

    for _ in range(warmup):
        step_device()                                  # result dropped at once: the timed steps reuse its memory
    profile = {"*": True}
    if os.environ.get("LEIDA_BENCH_DUMP"):
        profile["trace_active"] = True
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = nvl.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph_launches = 0
    step_marks = []
    stats = res = None
    for it in range(steps):
        res = None
        res = step_device(profile)
        graph_launches += res.batch.graph_launches
        b = res.batch
        stats = dict(run_passes=b.total_passes(), lockstep=b.lockstep_passes, n_runs=len(b._ks), exchange=b.exchange,
                     sum_k_passes=float(sum(b.k(i) * b.passes(i) for i in range(len(b._ks)))),
                     rechecked=b.rechecked_total, changed=b.changed_total)
        del b
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
    value = M_global * steps / (ms * 1e-3)

    parity = None
    if with_parity:
        R = stats["n_runs"]
        picks = sorted({0, R // 2, R - 1, R - 1 - wl["n_init"] // 2})      # K_min, a middle K, two restarts of K_max
        parity = fp64_parity(torch, res, wl, comm, row_lo, picks)
    res = None

    # roofline of the dominant kernel + the other kernels of the pass loop, from the event pairs recorded inside the
    # timed region (this rank)
    active_trace = profile.pop("active_after_pass", None)
    profile.pop("trace_active", None)
    kernel_ms = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in profile.items() if k != "*"}
    if os.environ.get("LEIDA_BENCH_DUMP") and rank == 0:
        # per-launch durations of the LAST timed step, kernel by kernel (profiles/: where the passes spend their time)
        per = {k: [round(a.elapsed_time(b), 4) for a, b in v[-(len(v) // steps):]] for k, v in profile.items() if k != "*"}
        with open(os.environ["LEIDA_BENCH_DUMP"], "w") as f:
            json.dump({"config": key, "n_gpus": world, "per_launch_ms": per, "active_after_pass": active_trace}, f)
    kernel_n = {k: len(v) for k, v in profile.items() if k != "*"}
    assign_ms = kernel_ms.get("assign", 0.0)
    n_assign = kernel_n.get("assign", 0)
    run_passes = stats["run_passes"]                    # per step (same every step: deterministic)
    N = wl["N"]
    npk = int(nvl.lib.leida_kmeans_tc_pitch(N))
    xp = (N + 31) // 32 * 32
    sum_k_passes = stats["sum_k_passes"]
    issued_flops = 3 * 2.0 * M_local * npk * sum_k_passes * steps          # bf16 x 3: X1*C1 + X1*C2 + X2*C1
    alg_flops = 2.0 * M_local * N * sum_k_passes * steps                    # 2*M*N*K per run per pass
    # kernel timed inside a long step (>= 0.25 s of back-to-back tensor work per step): sustained peak; else burst
    long_step = (ms / steps) >= 250.0
    tpeak, tsrc = tensor_peak(long_step)
    tburst = tensor_peak(False)[0]
    hpeak, hsrc = hbm_peak()
    sec = assign_ms * 1e-3
    ach = issued_flops / sec / 1e12 if sec > 0 else 0.0
    survey_bytes = (4.0 * M_local * N + 8.0 * M_local) * run_passes * steps
    actual_bytes = n_assign * (2 * 2.0 * M_local * npk) + 4.0 * M_local * run_passes * steps
    # ncu --set full of one full-width launch (all runs active), per workload; see profiles/
    traffic = TRAFFIC_NCU.get(key) if world == 1 else None
    roofline = {
        "kernel": "k_tc_assign (tcgen05 bf16x3 distance GEMM of all active runs + top-2 epilogue)",
        "bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak if tpeak else 0.0,
        "traffic": traffic,
        "traffic_launch": "dram bytes of ONE full-width launch (all runs active), ncu --set full, profiles/r02_ncu_pass0.txt; the SURVEY per-run figure for that launch is n_runs x (4*M*N + 8*M) -- 549 GB at config 3 -- i.e. no re-reads: one pass over the rows serves every run",
        "peak_source": tsrc, "peak_burst": tburst, "frac_of_burst": ach / tburst if tburst else 0.0,
        "launches": n_assign, "avg_launch_ms": assign_ms / max(n_assign, 1),
        "share_of_step": assign_ms / (ms if ms > 0 else 1.0),
        "flops_per_launch_issued": issued_flops / max(n_assign, 1),
        "algorithmic_tflops": alg_flops / sec / 1e12 if sec > 0 else 0.0,
        "algorithmic_frac": (alg_flops / sec / 1e12 / tpeak) if sec > 0 and tpeak else 0.0,
        "survey_metric": {"bound": "hbm", "achieved": survey_bytes / sec / 1e9 if sec > 0 else 0.0, "peak": hpeak,
                          "unit": "GB/s", "frac": (survey_bytes / sec / 1e9 / hpeak) if sec > 0 else 0.0,
                          "peak_source": hsrc,
                          "note": "SURVEY 8d per-run figure (4*M*N + 8*M bytes per run per Lloyd pass); one launch serves "
                                  "every active run from ONE read of the rows, so this exceeds the HBM peak by design",
                          "actual_bytes_per_launch_model": actual_bytes / max(n_assign, 1)},
        "note": "achieved = issued bf16 MMA work (3 MMAs per output for the bf16x3 split, K columns only, padded pitch) / "
                "sum of the launch durations; algorithmic_* counts 2*M*N*K once",
    }
    # the other kernels: achieved bytes on their own algorithmic traffic
    changed, rechecked = stats["changed"], stats["rechecked"]
    world_div = max(world, 1)
    models = {
        "recheck": ("k_recheck: fp64 re-evaluation of the undecided (row, run) pairs; bytes = pairs x (row + ~2.5 candidate "
                    "centroid rows) x 4*N", rechecked / world_div * steps * 3.5 * 4.0 * N),
        "apply": ("k_diff_apply: label diff + fixed-point moves; bytes = 8*M per active run-pass (two label rows) + 4*N per "
                  "changed row", (8.0 * M_local * run_passes + 4.0 * N * changed / world_div) * steps),
        "update": ("k_update: means, norms and bf16 packing of the active runs; bytes = (8 + 4 + 4) * K * pitch per run-pass",
                   16.0 * xp * sum_k_passes * steps),
        "distortion": ("k_distortion: fp64 distance of every row to its centroid, all runs; bytes = rows once + labels",
                       (4.0 * M_local * N + 4.0 * M_local * stats["n_runs"]) * steps),
    }
    bounds = {"recheck": "latency / fp32->fp64 conversion pipe (ncu: profiles/r02_ncu_pass0.txt)",
              "apply": "instruction issue in the dense passes (70 % of issue slots; the rows come from L2, hit rate 96 %)",
              "update": "latency (one dependent round trip per pass)",
              "distortion": "latency: 12 warps per SM (rows held in registers as doubles), per-(quad, run) reduction and "
                            "label handling; the warps of a CTA share the staged centroids, so the slowest quad paces them "
                            "(profiles/r02_ncu_distortion_piped.txt)"}
    kernels = {}
    for name, (what, nbytes) in models.items():
        if name in kernel_ms and kernel_ms[name] > 0:
            kernels[name] = {"what": what, "bound": bounds[name], "ms_per_step": kernel_ms[name] / steps,
                             "launches_per_step": kernel_n[name] / steps,
                             "achieved_GBs": nbytes / (kernel_ms[name] * 1e-3) / 1e9,
                             "hbm_frac": nbytes / (kernel_ms[name] * 1e-3) / 1e9 / hpeak}
    if "distortion" in kernels:
        kernels["distortion"]["fp64_tflops"] = 2.0 * M_local * N * stats["n_runs"] * steps / (kernel_ms["distortion"] * 1e-3) / 1e12
    if "stage12" in kernel_ms and kernel_ms["stage12"] > 0:
        s12_bytes = (8.0 * (hi - lo) * N * wl["T"] + 8.0 * M_local * N + 4.0 * M_local * xp) * steps
        kernels["stage12"] = {"what": "k_hilbert + k_eigvec_rank2 (fused stage 1+2); bytes = SURVEY 8d: 8*S*N*T in, 8*M*N + 4*M*NP out",
                              "bound": "latency (fp64 FFT in shared memory: barriers and input loads; ncu: profiles/r02_ncu_stage12.txt)",
                              "ms_per_step": kernel_ms["stage12"] / steps, "launches_per_step": kernel_n["stage12"] / steps,
                              "achieved_GBs": s12_bytes / (kernel_ms["stage12"] * 1e-3) / 1e9,
                              "hbm_frac": s12_bytes / (kernel_ms["stage12"] * 1e-3) / 1e9 / hpeak}
    for name in kernel_ms:
        if name not in models and name != "assign" and name not in kernels:
            kernels[name] = {"ms_per_step": kernel_ms[name] / steps, "launches_per_step": kernel_n[name] / steps}
    roofline["kernels"] = kernels
    roofline["other_kernels_ms_per_step"] = {k: v / steps for k, v in kernel_ms.items() if k != "assign"}

    e2e = None
    if with_e2e:
        for _ in range(2):
            step_e2e()                                     # warm-up (pinned staging buffers, allocator); result dropped
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = steps
        f0.record()
        t_wall = time.perf_counter()
        h2d = d2h = 0
        phases = {}
        for _ in range(n_e2e):
            m = step_e2e()
            h2d, d2h = int(m.h2d_bytes), int(m.d2h_bytes)
            for kk, vv in getattr(m, "timings_", {}).items():
                phases[kk] = phases.get(kk, 0.0) + vv / n_e2e
            del m
        f1.record()
        barrier()
        t_wall = time.perf_counter() - t_wall
        e2e_ms = max(f0.elapsed_time(f1), 1e3 * t_wall)
        te = torch.tensor([e2e_ms, float(h2d), float(d2h)], dtype=torch.float64, device=dev)
        tb = te[1:].clone()
        if world > 1:
            dist.all_reduce(te[:1], op=dist.ReduceOp.MAX)
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        e2e = {"value": M_global * n_e2e / (float(te[0].item()) * 1e-3), "unit": "volumes/s",
               "h2d_bytes_per_step": int(tb[0].item()), "d2h_bytes_per_step": int(tb[1].item()), "steps": n_e2e,
               "ms_per_step": float(te[0].item()) / n_e2e,
               "host_phases_ms_rank0": {kk: round(vv, 1) for kk, vv in phases.items()},
               "call": "Leida.from_arrays(ts, classes).fit_predict(..., scores=False, stats=False) + the three metric tables "
                       "of every K: the metric is phase->eigvec->k-means->metrics (SURVEY 8d); quality scores and "
                       "permutation tests are separate stages"}
    run_stats = {"subjects_on_rank0": hi - lo, "rows_total": M_global, "kmeans_run_passes_per_step": run_passes,
                 "lockstep_passes_per_step": stats["lockstep"], "fp64_rechecks_per_step": rechecked,
                 "label_changes_per_step": changed, "per_step_ms": [round(v, 2) for v in per_step_ms],
                 "input_generation_s": round(t_gen, 1), "parallelism": f"subjects sharded over {world} GPU(s)",
                 "exchange": stats["exchange"]}
    return dict(value=value, ms_per_step=ms / steps, launches=int(ln.item()), roofline=roofline, e2e=e2e, parity=parity,
                clocks=clocks, run_stats=run_stats, stats=stats)


# dram__bytes_read.sum + dram__bytes_write.sum of one full-width k_tc_assign launch (ncu --set full), per workload
TRAFFIC_NCU = {"config2": 153.8e6,          # profiles/r01_k_tc_assign_ncu_full.txt, capture F
               "config3": 3.603e9}          # profiles/r02_ncu_pass0.txt: 2.245 GB read (the bf16x2 rows once) + 1.358 GB written (labels)


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
    key, wl = args.config, WORKLOADS[args.config]
    r = measure(torch, dist, args, key, wl, args.steps, args.warmup, with_e2e=not args.no_e2e)
    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": r["value"], "unit": "volumes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": static_config(key, wl),
            "run_stats": r["run_stats"], "roofline": r["roofline"], "e2e": r["e2e"], "parity": r["parity"],
            "gpu_launches": r["launches"], "clocks": r["clocks"],
        }
    if world == 1 and key == DEFAULT_CONFIG and not args.no_secondary:
        k2, w2 = "config2", WORKLOADS["config2"]
        torch.cuda.empty_cache()
        r2 = measure(torch, dist, args, k2, w2, min(args.steps, 5), 3, with_e2e=True, with_parity=True)
        out["secondary"] = {"config": static_config(k2, w2), "value": r2["value"], "unit": "volumes/s",
                            "ms_per_step": r2["ms_per_step"], "e2e": r2["e2e"], "parity": r2["parity"],
                            "run_stats": r2["run_stats"],
                            "roofline": {k: r2["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac",
                                                                        "avg_launch_ms", "share_of_step")}}
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample_subjects, n_init_sample = CPU_SAMPLE[key]
        times = run_cpu_sample(wl, sample_subjects, n_init_sample, 1, 0, cores)
        full_ppr = r["stats"]["run_passes"] / float(r["stats"]["n_runs"])
        cv, _, cv_iter = cpu_value(times, wl, n_init_sample, full_ppr)
        out["cpu_baseline"] = {
            "value": cv, "unit": "volumes/s", "cores": cores, "kind": "port", "extrapolated": True,
            "value_iteration_matched": cv_iter,
            "sample": cpu_sample_text(key, wl, sample_subjects, n_init_sample, cores, times[0]["wall"]) +
            f"; value_iteration_matched additionally scales the k-means time to the {full_ppr:.1f} passes per run the "
            f"full-size GPU run needed (the sample converged in {times[0]['lloyd_iters'] / times[0]['runs']:.1f} iterations per run)"}
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
    ap.add_argument("--config", default=DEFAULT_CONFIG, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the config-2 secondary measurement")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end (host API) leg: profiling runs only")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return native_arm(args)


if __name__ == "__main__":
    sys.exit(main())
