"""Host side of stage 3: KMeansLeida / identify_states over the CUDA lock-step k-means engine.

Reference: pyleida/clustering/_clustering.py:24-374.  What stays on the host is exactly what the
reference does with numpy's *global legacy RNG* and tiny arrays: drawing the initial rows
(:103-105), choosing the best restart (:139-146) and the relabelling permutation (:234-247).
Everything that touches the (M, N) data runs in libleida_b200.so.
"""
import os
import warnings

import numpy as np

from .. import _native as nv
from ..parallel import Comm, PeerExchange, gather_seed_rows
from ..signal_tools.phase import padded_pitch

_KP_BUCKETS = (4, 8, 12, 16, 20, 24, 32, 40, 48, 64)
# 'tc': tcgen05 tensor-core assignment (production path); 'simt': fp32 CUDA-core kernel kept for
# cross-checking the two implementations against each other in the tests.
DEFAULT_ASSIGN = os.environ.get("LEIDA_KMEANS_ASSIGN", "tc")
# Replaying the pass as a CUDA graph is supported but off by default: the eager loop needs ~50 us of host
# time per pass and is GPU-bound, while capturing costs tens of ms per run() (measured: tools/runcost.py).
USE_CUDA_GRAPHS = os.environ.get("LEIDA_CUDA_GRAPHS", "0") != "0"


def _bucket(k):
    for b in _KP_BUCKETS:
        if k <= b:
            return b
    raise ValueError(f"k must be <= {nv.KM_MAX_K}")


class KMeansEngine:
    """Device-resident rows + lock-step execution of many (K, initial rows) runs.

    X32: CUDA float32 tensor (M_local, pitch), pitch = padded_pitch(N), pad columns zero (the
    layout leading_eigenvectors() produces), or any (M, N) float array (packed here).
    comm / row_lo / m_global describe row sharding over ranks (defaults: single GPU).
    """

    def __init__(self, X, n_features=None, comm=None, row_lo=0, m_global=None):
        torch = nv.require_cuda()
        self.torch = torch
        if isinstance(X, np.ndarray) or not torch.is_tensor(X):
            X = torch.from_numpy(np.ascontiguousarray(np.asarray(X), dtype=np.float32)).cuda()
        if not X.is_cuda:
            X = X.cuda()
        if X.dim() != 2:
            raise ValueError("expected a 2D (n_samples, n_features) array")
        N = int(n_features) if n_features is not None else int(X.shape[1])
        xp = padded_pitch(N)
        if X.dtype != torch.float32 or X.shape[1] != xp or not X.is_contiguous():
            src = X.to(torch.float32).contiguous()
            packed = torch.empty((src.shape[0], xp), dtype=torch.float32, device=src.device)
            nv.call("leida_kmeans_pack_rows", nv.ptr(src), src.shape[0], N, src.shape[1], nv.ptr(packed), xp, nv.stream())
            X = packed
        self.N, self.XP = N, xp
        self.M = int(X.shape[0])
        self.comm = comm if comm is not None else Comm()
        self.row_lo = int(row_lo)
        self.m_global = int(m_global) if m_global is not None else self.M
        # Range of the fixed-point centroid sums (int64, 2^40 per unit): exact while |x| <= 1 and a cluster holds fewer
        # than 2^22 rows.  LEiDA eigenvectors are unit vectors; any other data is brought into range by ONE global
        # power-of-two factor (exact in float32; cosine assignment and distortion do not depend on it, and the
        # centroids are scaled back exactly where they leave the engine).  Elements below 2^-40 of the largest
        # magnitude do not contribute to the sums.
        if self.m_global >= (1 << 22):
            raise ValueError("KMeansEngine supports fewer than 2^22 rows in total (int64 fixed-point centroid sums)")
        amax = 0.0
        if self.M > 0:
            mn, mx = X.aminmax()                          # no (M, pitch) temporary
            amax = max(-float(mn), float(mx))
        if self.comm.world > 1:
            t = torch.tensor([amax], dtype=torch.float64, device=X.device if self.comm.backend == "nccl" else "cpu")
            amax = float(self.comm.all_reduce_max_(t).item())
        if not np.isfinite(amax):
            raise ValueError("rows must be finite")
        self.scale = 1.0
        if amax > 1.0 or 0.0 < amax < 2.0 ** -8:
            self.scale = float(2.0 ** np.ceil(np.log2(amax)))
            X = X * (1.0 / self.scale)
        self.X = X
        self.xnorm = torch.empty(max(self.M, 1), dtype=torch.float64, device=X.device)
        nv.call("leida_kmeans_row_norms", nv.ptr(X), self.M, N, xp, nv.ptr(self.xnorm), nv.stream())
        # |x| <= 1 everywhere lets leida_kmeans_tc_apply quantise on the FMA pipe
        self.unit_rows = self.M > 0

    # ------------------------------------------------------------------------------------------
    def _tc_operands(self):
        """bf16 x 2 split of the rows for the tensor-core path (made once per data set)."""
        if getattr(self, "_x12", None) is None:
            torch = self.torch
            npk = int(nv.lib.leida_kmeans_tc_pitch(self.N))
            x1 = torch.empty((self.M, npk), dtype=torch.bfloat16, device=self.X.device)
            x2 = torch.empty((self.M, npk), dtype=torch.bfloat16, device=self.X.device)
            nv.call("leida_kmeans_tc_split_rows", nv.ptr(self.X), self.M, self.N, self.XP, nv.ptr(x1), nv.ptr(x2), nv.stream())
            self._x12 = (x1, x2, npk)
        return self._x12

    def _capture(self, do_assign, do_rest, split):
        """CUDA graphs of one lock-step pass: [assign + rest], or [assign], [rest] when the assign
        kernel has to be timed on its own."""
        torch = self.torch
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graphs = []
        with torch.cuda.stream(side):
            for parts in ((do_assign,), (do_rest,)) if split else ((do_assign, do_rest),):
                g = torch.cuda.CUDAGraph()
                g.capture_begin()
                for fn in parts:
                    fn()
                g.capture_end()
                graphs.append(g)
        torch.cuda.current_stream().wait_stream(side)
        return graphs

    def run(self, runs, n_iter=1000, centroid_update="exact", poll_every=1, profile=None, assign=None, metric="cosine"):
        """runs: list of (k, init_rows) with init_rows = k global row indices.  Returns BatchResult.
        profile: optional dict; for every kernel name present as a key ("assign", "recheck", "apply",
        "update", "plan"; or all of them with {"*": True}) the (start, stop) CUDA event pairs bracketing
        each launch are appended to profile[name] (bench.py's live roofline measurement)."""
        torch = self.torch
        if centroid_update not in ("exact", "reference"):
            raise ValueError("centroid_update must be 'exact' or 'reference'")
        if metric not in ("cosine", "euclidean"):
            raise ValueError("'metric' must be either 'cosine' or 'euclidean'")
        metric_id = 1 if metric == "euclidean" else 0
        assign = assign or DEFAULT_ASSIGN
        if metric_id:
            assign = "simt"                      # the tensor-core flow is cosine only
        if assign not in ("tc", "simt"):
            raise ValueError("assign must be 'tc' (tensor cores) or 'simt'")
        if centroid_update == "reference" and self.comm.world > 1:
            raise ValueError("centroid_update='reference' (sequential float32 sums) cannot be sharded over GPUs")
        if self.M < 1:
            raise ValueError("every rank needs at least one row")
        dev = self.X.device
        order = sorted(range(len(runs)), key=lambda i: (runs[i][0], i))
        ks = np.array([runs[i][0] for i in order], dtype=np.int32)
        if ks.min() < 2 or ks.max() > nv.KM_MAX_K:
            raise ValueError(f"k must be in [2, {nv.KM_MAX_K}]")
        crow0 = np.zeros(len(order), dtype=np.int32)
        crow0[1:] = np.cumsum(ks[:-1])
        CR = int(ks.sum())
        R = len(order)
        seeds_global = np.concatenate([np.asarray(runs[i][1], dtype=np.int64).reshape(-1) for i in order])
        if seeds_global.size != CR:
            raise ValueError("each run needs exactly k initial rows")
        if seeds_global.min() < 0 or seeds_global.max() >= self.m_global:
            raise ValueError("initial row index out of range")

        run_k = torch.from_numpy(ks).to(dev)
        run_c0 = torch.from_numpy(crow0).to(dev)
        AP = self.XP + 8
        cent = torch.empty((CR, self.XP), dtype=torch.float32, device=dev)
        cnorm = torch.empty(CR, dtype=torch.float64, device=dev)
        # acc and the run state live in ONE int64 buffer: with several ranks a pass needs a single copy and a
        # single all-reduce (only the reducible state words are read from the reduced copy)
        n_acc = CR * AP
        red_local = torch.empty(n_acc + R * nv.STATE_WORDS, dtype=torch.int64, device=dev)
        acc = red_local[:n_acc].view(CR, AP)
        state = red_local[n_acc:].view(R, nv.STATE_WORDS)
        labels = torch.empty((R, self.M), dtype=torch.int32, device=dev)
        # (run,row) pairs queued for the fp64 re-check per pass: 32-byte entries; a pass needs a fraction of a percent of
        # R*M, and a full queue only costs time (the re-check then scans for the unlisted marks)
        cap = int(min(max(R * self.M // 64, 1 << 16), 1 << 23))
        cap = int(os.environ.get("LEIDA_KMEANS_WORKLIST_CAP", cap))   # tests shrink it to exercise the overflow path
        worklist = torch.empty(8 * cap + 8, dtype=torch.int32, device=dev)
        worklist[:8].zero_()
        n_active = torch.zeros(8, dtype=torch.int32, device=dev)
        mailbox = torch.zeros(4, dtype=torch.int32, pin_memory=True)   # written by the last CTA of leida_kmeans_update
        mb64 = mailbox.numpy().view(np.int64)          # one packed word: active | passes done << 16 | first stop << 40

        def mb_read():
            v = int(mb64[0])
            return v & 0xFFFF, (v >> 16) & 0xFFFFFF, (v >> 40) & 0xFFFFFF
        st = nv.stream()

        peers = None
        if self.comm.world > 1:
            seed_rows = gather_seed_rows(self.X, self.N, seeds_global, self.row_lo, self.comm)
            init_idx = torch.arange(CR, dtype=torch.int64, device=dev)
            nv.call("leida_kmeans_init", nv.ptr(seed_rows), self.XP, nv.ptr(init_idx), R, nv.ptr(run_k), nv.ptr(run_c0),
                    self.N, self.XP, nv.ptr(cent), nv.ptr(cnorm), nv.ptr(acc), nv.ptr(labels), self.M, nv.ptr(state), st)
            red_global = torch.empty_like(red_local)
            acc_red = red_global[:n_acc].view(CR, AP)
            state_red = red_global[n_acc:].view(R, nv.STATE_WORDS)
            # per-pass exchange: peer-memory reads fused into leida_kmeans_update_peers when the ranks can map
            # each other's buffers, else copy + NCCL all-reduce into red_global
            peers = PeerExchange.get(self.comm, red_local.numel(), dev) if centroid_update == "exact" else None
        else:
            init_idx = torch.from_numpy(seeds_global).to(dev)
            nv.call("leida_kmeans_init", nv.ptr(self.X), self.XP, nv.ptr(init_idx), R, nv.ptr(run_k), nv.ptr(run_c0),
                    self.N, self.XP, nv.ptr(cent), nv.ptr(cnorm), nv.ptr(acc), nv.ptr(labels), self.M, nv.ptr(state), st)
            acc_red, state_red = acc, state

        capturing = [False]      # True while _capture() records the launches on its side stream

        def timed(name, fn):
            if profile is None or not (profile.get("*") or name in profile):
                return fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            profile.setdefault(name, []).append((e0, e1))

        if assign == "tc" and R > 4096:
            # the tensor-core packing plan holds at most 4096 runs per batch (kmeans_tc.cu: PLAN_MAX_RUNS)
            warnings.warn(f"{R} runs in one batch exceed the tensor-core plan's 4096: using the CUDA-core assignment")
            assign = "simt"
        tc = assign == "tc"
        # trust threshold of the tensor-core argmax: the library's measured bound, optionally scaled (experiments only)
        tau = float(nv.lib.leida_kmeans_tc_tau(self.N)) * float(os.environ.get("LEIDA_TC_TAU_SCALE", "1"))
        if tc:
            x1, x2, npk = self._tc_operands()
            # packed columns: every run owns a slot of read_width(K) <= 2K columns; tiles hold one slot width each,
            # so up to one partly filled tile per width class
            bn = int(nv.lib.leida_kmeans_tc_tile_columns())
            c_rows = (4 * CR + bn * 17 + bn - 1) // bn * bn
            c1 = torch.zeros((c_rows, npk), dtype=torch.bfloat16, device=dev)
            c2 = torch.zeros((c_rows, npk), dtype=torch.bfloat16, device=dev)
            labels_new = torch.empty((R, self.M), dtype=torch.int32, device=dev)
            plan = torch.zeros(2, dtype=torch.int32, device=dev)
            tile_first = torch.zeros(c_rows // 128, dtype=torch.int32, device=dev)
            tile_count = torch.zeros(c_rows // 128, dtype=torch.int32, device=dev)
            prun = torch.zeros(R, dtype=torch.int32, device=dev)
            pcol = torch.zeros(R, dtype=torch.int32, device=dev)
            run_pcol = torch.zeros(R, dtype=torch.int32, device=dev)

        def tc_plan():
            nv.call("leida_kmeans_tc_plan", R, nv.ptr(run_k), nv.ptr(run_c0), nv.ptr(state), nv.ptr(cent), nv.ptr(cnorm),
                    self.N, self.XP, nv.ptr(plan), nv.ptr(tile_first), nv.ptr(tile_count), nv.ptr(prun), nv.ptr(pcol),
                    nv.ptr(run_pcol), nv.ptr(c1), nv.ptr(c2), 1, nv.stream())

        # groups of consecutive (K-sorted) runs sharing a register tile (SIMT flow only)
        groups, g0 = [], 0
        for i in range(1, R + 1):
            if i == R or _bucket(int(ks[i])) != _bucket(int(ks[g0])):
                groups.append((g0, i, int(ks[g0:i].max())))
                g0 = i

        # The launches of a pass never change their arguments (except the apply grid bound after a re-plan), and in
        # the late passes the GPU needs less time for a pass than Python needs to marshal four calls: bind them once.
        fast = {}

        def do_assign():
            stq = st if not capturing[0] else nv.stream()
            if tc:
                f = fast.get(("assign", stq))
                if f is None:
                    f = fast[("assign", stq)] = nv.prepared(
                        "leida_kmeans_tc_assign", nv.ptr(x1), nv.ptr(x2), nv.ptr(self.xnorm), self.M, self.N, nv.ptr(c1), nv.ptr(c2),
                        c_rows, nv.ptr(plan), nv.ptr(tile_first), nv.ptr(tile_count), nv.ptr(prun), nv.ptr(labels_new),
                        nv.ptr(worklist), cap, tau, stq)
                f()
                return
            for (a, b, kmax) in groups:
                nv.call("leida_kmeans_assign", nv.ptr(self.X), nv.ptr(self.xnorm), self.M, self.N, self.XP, b - a, kmax,
                        run_k.data_ptr() + 4 * a, run_c0.data_ptr() + 4 * a, nv.ptr(cent), nv.ptr(cnorm), nv.ptr(acc),
                        labels.data_ptr() + 4 * a * self.M, state.data_ptr() + 8 * nv.STATE_WORDS * a,
                        nv.ptr(worklist), cap, metric_id, stq)
                nv.call("leida_kmeans_recheck", nv.ptr(self.X), nv.ptr(self.xnorm), self.M, self.N, self.XP, b - a,
                        run_k.data_ptr() + 4 * a, run_c0.data_ptr() + 4 * a, nv.ptr(cent), nv.ptr(cnorm), nv.ptr(acc),
                        labels.data_ptr() + 4 * a * self.M, state.data_ptr() + 8 * nv.STATE_WORDS * a,
                        nv.ptr(worklist), cap, None, metric_id, stq)

        def do_rest():
            stq = st if not capturing[0] else nv.stream()
            if tc:
                f = fast.get(("recheck", stq))
                if f is None:
                    f = fast[("recheck", stq)] = nv.prepared(
                        "leida_kmeans_recheck", nv.ptr(self.X), nv.ptr(self.xnorm), self.M, self.N, self.XP, R, nv.ptr(run_k),
                        nv.ptr(run_c0), nv.ptr(cent), nv.ptr(cnorm), nv.ptr(acc), nv.ptr(labels), nv.ptr(state),
                        nv.ptr(worklist), cap, nv.ptr(labels_new), 0, stq)
                timed("recheck", f)
                f = fast.get(("apply", stq, grid_bound[0]))
                if f is None:
                    f = fast[("apply", stq, grid_bound[0])] = nv.prepared(
                        "leida_kmeans_tc_apply", nv.ptr(self.X), self.M, self.N, self.XP, int(self.unit_rows), int(grid_bound[0]),
                        nv.ptr(plan), nv.ptr(prun), nv.ptr(run_k), nv.ptr(run_c0), nv.ptr(labels), nv.ptr(labels_new), nv.ptr(acc),
                        nv.ptr(state), stq)
                timed("apply", f)
            if fused and (self.comm.world == 1 or peers is not None):
                # one kernel: [publish the run's words, raise / await the per-run flags, read sums over ranks,] means,
                # norms and packed operands from registers
                hkey = peers.half if peers is not None else 0
                f = fast.get(("fused", stq, hkey))
                if f is None:
                    if peers is not None:
                        half_ptr, ptrs, mc = peers.next_half()
                        xargs = (half_ptr, nv.ptr(ptrs), mc or None, nv.ptr(peers._flag_ptrs), peers.flags_local,
                                 peers.FLAG_PITCH, peers.world, peers.rank, peers.ctr.data_ptr() + 8)
                    else:
                        xargs = (None, None, None, None, None, 0, 1, 0, None)
                    f = fast[("fused", stq, hkey)] = nv.prepared(
                        "leida_kmeans_update_fused", R, nv.ptr(run_k), nv.ptr(run_c0), self.N, self.XP, nv.ptr(red_local),
                        n_acc, *xargs, nv.ptr(cent), nv.ptr(cnorm), nv.ptr(state), int(n_iter), nv.ptr(n_active),
                        nv.ptr(run_pcol) if tc else None, nv.ptr(c1) if tc else None, nv.ptr(c2) if tc else None,
                        mailbox.data_ptr(), stq)
                elif peers is not None:
                    peers.next_half()
                timed("update", f)
                if centroid_update == "reference":
                    nv.call("leida_kmeans_means_f32_sequential", nv.ptr(self.X), self.M, self.N, self.XP, R, nv.ptr(run_k),
                            nv.ptr(run_c0), nv.ptr(labels), nv.ptr(state), nv.ptr(cent), nv.ptr(cnorm), stq)
                    if tc:
                        tc_plan()                          # the float32 sequential means replace the centroids: re-pack
                return
            if self.comm.world > 1 and peers is not None:
                f = fast.get(("xchg", stq, peers.half))
                if f is None:
                    key = ("xchg", stq, peers.half)
                    half_ptr, ptrs, mc = peers.next_half()
                    pub = nv.prepared("leida_kmeans_exchange_publish", R, nv.ptr(run_k), nv.ptr(run_c0), self.XP,
                                      nv.ptr(red_local), n_acc, half_ptr, nv.ptr(peers._flag_ptrs), peers.world, peers.rank,
                                      peers.ctr.data_ptr(), nv.ptr(peers.ticket), stq)
                    upd = nv.prepared("leida_kmeans_update_peers", R, nv.ptr(run_k), nv.ptr(run_c0), self.N, self.XP,
                                      nv.ptr(ptrs), peers.world, mc or None, n_acc, peers.flags_local,
                                      peers.ctr.data_ptr() + 8, nv.ptr(cent), nv.ptr(cnorm), nv.ptr(state), int(n_iter),
                                      nv.ptr(n_active), nv.ptr(run_pcol) if tc else None, nv.ptr(c1) if tc else None,
                                      nv.ptr(c2) if tc else None, mailbox.data_ptr(), stq)
                    f = fast[key] = (pub, upd)
                else:
                    peers.next_half()
                timed("publish", f[0])
                timed("update", f[1])
                return
            if self.comm.world > 1:
                red_global.copy_(red_local)
                self.comm.all_reduce_sum_(red_global)
            f = fast.get(("update", stq))
            if f is None:
                f = fast[("update", stq)] = nv.prepared(
                    "leida_kmeans_update", R, nv.ptr(run_k), nv.ptr(run_c0), self.N, self.XP, nv.ptr(acc_red),
                    nv.ptr(state_red), nv.ptr(cent), nv.ptr(cnorm), nv.ptr(state), int(n_iter), nv.ptr(n_active),
                    nv.ptr(run_pcol) if tc else None, nv.ptr(c1) if tc else None, nv.ptr(c2) if tc else None,
                    mailbox.data_ptr(), stq)
            timed("update", f)
            if centroid_update == "reference":
                nv.call("leida_kmeans_means_f32_sequential", nv.ptr(self.X), self.M, self.N, self.XP, R, nv.ptr(run_k),
                        nv.ptr(run_c0), nv.ptr(labels), nv.ptr(state), nv.ptr(cent), nv.ptr(cnorm), stq)
                if tc:
                    tc_plan()                          # the float32 sequential means replace the centroids: re-pack

        # update + exchange in one kernel (rows up to 1024 columns, at most FLAG_PITCH runs when exchanging)
        fused = self.XP <= 1024 and os.environ.get("LEIDA_UPDATE_LEGACY") is None and \
            (peers is None or R <= PeerExchange.FLAG_PITCH)
        launches_per_pass = (3 if tc else 2 * len(groups)) + 1 + (2 if centroid_update == "reference" else 0)
        if tc:
            tc_plan()
        # Lock-step passes.  The first pass runs eagerly; the following ones replay CUDA graphs of the
        # (static) launch sequence.  The host never synchronises inside the loop: the last CTA of
        # leida_kmeans_update stores {runs still active, passes completed} in a pinned mailbox, which
        # the host reads to stop, to bound its run-ahead, and to re-pack the tensor-core operands once
        # enough runs have stopped (the active count never increases, so a stale value is a valid
        # upper bound; passes issued after every run has stopped are no-ops).
        use_graph = tc and centroid_update == "exact" and USE_CUDA_GRAPHS and self.comm.world == 1
        time_assign = profile is not None and (profile.get("*") or "assign" in profile)
        eager_rest = profile is not None and any(k in profile for k in ("*", "recheck", "apply", "update"))
        max_lag = 6
        mb64[:] = 0
        packed_bound = R
        active_bound = R
        grid_bound = [R]        # upper bound of the packed-run count, sizes the apply grid (eager launches only)
        issued = 0
        graphs = None
        act_hist = torch.zeros(2048, dtype=torch.int32, device=dev) if (profile is not None and profile.get("trace_active")) else None
        while True:
            if tc and mb_read()[2] == 0 and 0 < active_bound <= 0.9 * packed_bound:
                tc_plan()                              # re-pack: only the runs still active keep tensor-core columns
                packed_bound = active_bound
                if not use_graph:
                    grid_bound[0] = max(1, packed_bound)
            if use_graph and issued == 1 and not eager_rest:
                capturing[0] = True
                graphs = self._capture(do_assign, do_rest, split=time_assign)
                capturing[0] = False
            if graphs is None:
                timed("assign", do_assign)
                do_rest()
            elif len(graphs) == 2:
                timed("assign", graphs[0].replay)
                graphs[1].replay()
            else:
                graphs[0].replay()
            if act_hist is not None and issued < act_hist.numel():
                act_hist[issued:issued + 1].copy_(n_active[2:3])       # runs still active after this pass (profiling only)
            issued += 1
            while issued - mb_read()[1] > max_lag:
                pass
            n_act, n_done, first_stop = mb_read()          # one word: the three fields belong to the same pass
            if n_done >= 1:
                active_bound = min(active_bound, n_act)
            # Stop rule, identical on every rank (the collectives inside the passes must match): the
            # device records the pass p0 after which no run was active; because the host checks after
            # every issue and is never more than max_lag passes ahead, issued <= p0 + max_lag when p0
            # becomes visible, and every rank issues exactly p0 + max_lag passes.
            if first_stop > 0 and issued >= first_stop + max_lag:
                break
            if issued > n_iter + 2 + 2 * max_lag:
                raise nv.LeidaCudaError("k-means did not terminate (internal error)")
        passes = issued
        if act_hist is not None:
            profile["active_after_pass"] = act_hist[:min(issued, 2048)].cpu().tolist()
        launches = issued * launches_per_pass
        graph_launches = (issued - 1) * launches_per_pass if graphs is not None else 0   # not seen by the C-side counter
        timed("distortion", lambda: nv.call(
            "leida_kmeans_distortion", nv.ptr(self.X), nv.ptr(self.xnorm), self.M, self.N, self.XP, R, int(ks.max()), nv.ptr(run_k),
            nv.ptr(run_c0), nv.ptr(cent), nv.ptr(cnorm), nv.ptr(labels), nv.ptr(state), metric_id, st))
        launches += 2
        if self.comm.world > 1:
            red_global.copy_(red_local)
            self.comm.all_reduce_sum_(red_global)
            dist_words = state_red[:, [nv.ST_DIST_HI, nv.ST_DIST_LO]].cpu().numpy()
        else:
            dist_words = state[:, [nv.ST_DIST_HI, nv.ST_DIST_LO]].cpu().numpy()
        st_host = state.cpu().numpy()
        counts = acc_red[:, self.XP].cpu().numpy()
        hi = dist_words[:, 0].astype(np.float64)
        lo = dist_words[:, 1].astype(np.float64)
        inertia = hi * 2.0 ** -40 + lo * 2.0 ** -80
        if metric_id:
            inertia = inertia * self.scale ** 2          # squared Euclidean distances of the range-scaled rows
        inertia[dist_words[:, 0] >= (1 << 62)] = np.nan
        empty = st_host[:, nv.ST_EMPTY].copy()
        inertia[empty > 0] = np.nan
        inv = np.empty(R, dtype=np.int64)
        inv[np.asarray(order)] = np.arange(R)
        out = BatchResult(self, inv, ks, crow0, cent, labels, inertia, st_host[:, nv.ST_PASSES].copy(), empty, counts,
                          launches, passes)
        # largest tensor-core margin (fraction of the trust threshold) at which the fp64 re-check overruled the tensor-core
        # argmax: the error bound is sound while this stays well below 1 (see kmeans_tc.cu: default_tau)
        out.tau = tau
        out.tau_watch = float(st_host[:, nv.ST_TAU_WATCH].astype(np.int32).view(np.float32).max()) if tc else 0.0
        if out.tau_watch > 0.5:
            warnings.warn(f"tensor-core trust threshold is thin: an argmax with margin {out.tau_watch:.2f} x threshold was "
                          "overruled in fp64; results are still exact for every re-checked pair, but rerun with "
                          "LEIDA_TC_TAU_SCALE=2 (or LEIDA_KMEANS_ASSIGN=simt) and report this data set")
        out.rechecked_total = int(st_host[:, nv.ST_RECHECK_TOTAL].sum())
        out.changed_total = int(st_host[:, nv.ST_CHANGED_TOTAL].sum())
        out.assign = assign
        out.metric = metric
        out.exchange = ("none" if self.comm.world == 1 else "nccl all-reduce" if peers is None else
                        ("symmetric memory, multimem.ld_reduce" if peers.multicast else "symmetric memory, loads per rank") +
                        (", fused into the update kernel (per-run flags)" if fused else ", publish + update kernels"))
        out.graph_launches = graph_launches
        out._acc = acc_red                    # exact fixed-point member sums (device), for centers_exact()
        return out

    def predict(self, centers, metric="cosine"):
        """argmin distance to `centers` (K, N) for every local row -> CUDA int32 (M,)."""
        torch = self.torch
        c = torch.as_tensor(np.ascontiguousarray(centers, dtype=np.float32)).to(self.X.device)
        if metric == "euclidean" and self.scale != 1.0:
            c = c * (1.0 / self.scale)                   # the rows are held range-scaled
        out = torch.empty(self.M, dtype=torch.int32, device=self.X.device)
        nv.call("leida_kmeans_transform_f64", nv.ptr(self.X), self.M, self.N, self.XP, nv.ptr(c), c.shape[0], c.shape[1],
                None, nv.ptr(out), 1 if metric == "euclidean" else 0, nv.stream())
        return out


class BatchResult:
    """Results of KMeansEngine.run, indexed by the caller's run order."""

    def __init__(self, engine, inv, ks, crow0, cent, labels, inertia, passes, empty, counts, launches, lockstep_passes):
        self.engine, self._inv = engine, inv
        self._ks, self._crow0, self._cent, self._labels = ks, crow0, cent, labels
        self._inertia, self._passes, self._empty, self._counts = inertia, passes, empty, counts
        self.launches = launches
        self.lockstep_passes = lockstep_passes

    def k(self, i): return int(self._ks[self._inv[i]])
    def inertia(self, i): return float(self._inertia[self._inv[i]])
    def passes(self, i): return int(self._passes[self._inv[i]])
    def empty_cluster(self, i): return int(self._empty[self._inv[i]])

    def total_passes(self):
        return int(self._passes.sum())

    def counts(self, i):
        j = self._inv[i]
        return self._counts[self._crow0[j]:self._crow0[j] + self._ks[j]].astype(np.int64)

    def centers_device(self, i):
        j = self._inv[i]
        c = self._cent[int(self._crow0[j]):int(self._crow0[j] + self._ks[j]), :self.engine.N]
        return c if self.engine.scale == 1.0 else c * self.engine.scale

    def labels_device(self, i):
        return self._labels[int(self._inv[i])]

    def centers_exact(self, i):
        """(K, N) float64 member means of run i from the fixed-point sums (empty clusters -> NaN)."""
        j = self._inv[i]
        a, b = int(self._crow0[j]), int(self._crow0[j] + self._ks[j])
        sums = self._acc[a:b, :self.engine.N].cpu().numpy().astype(np.float64) * 2.0 ** -nv.LEIDA_KM_FRAC_BITS
        with np.errstate(divide="ignore", invalid="ignore"):
            return sums / self._counts[a:b].astype(np.float64)[:, None] * self.engine.scale


def _remap_permutation(counts):
    """KMeansLeida._remap_labels (clustering/_clustering.py:234-247) from the cluster sizes.

    Returns (lut, order): new_label = lut[old_label]; centers_new = centers[order].  Uses the same
    numpy calls as the reference so ties between equally populated clusters resolve identically.
    """
    counts = np.asarray(counts, dtype=np.int64)
    label = np.nonzero(counts)[0]                     # np.unique(labels_) of :242
    freq = counts[label]                              # its return_counts
    sorted_idxs = np.argsort(freq)[::-1]              # :243
    if len(label) != len(counts):
        # The reference indexes the *unique* list by position (:244-247) and breaks when a cluster
        # is empty; here: populated clusters by descending size, empty ones last.
        warnings.warn("a cluster ended empty; relabelling by descending size with empty clusters last")
        order = np.concatenate([label[sorted_idxs], np.nonzero(counts == 0)[0]])
        lut = np.empty(len(counts), dtype=np.int32)
        lut[order] = np.arange(len(counts), dtype=np.int32)
        return lut, order
    lut = np.empty(len(counts), dtype=np.int32)
    lut[sorted_idxs] = label.astype(np.int32)         # dct = {sorted_idxs[i]: label[i]}  (:244)
    return lut, sorted_idxs


def _select_best(inertias):
    """Restart selection of clustering/_clustering.py:139-146: first run, replaced only by a
    strictly smaller distortion.  Deviation: runs that hit an empty cluster (inertia NaN -- the
    reference would carry NaN centroids) are skipped with a warning instead of winning by default."""
    best = None
    for i, d in enumerate(inertias):
        if np.isnan(d):
            continue
        if best is None or d < inertias[best]:
            best = i
    if best is None:
        raise RuntimeError("every k-means restart ended with an empty cluster")
    if any(np.isnan(d) for d in inertias):
        warnings.warn("some k-means restarts hit an empty cluster and were discarded")
    return best


def _draw_inits(M, k, n_init, random_state):
    """Initial rows exactly as the reference draws them (clustering/_clustering.py:101-105):
    re-seeding before every restart when random_state is an int (=> identical restarts)."""
    out = []
    for _ in range(n_init):
        if random_state is not None:
            np.random.seed(random_state)
        out.append(np.random.choice(M, k, replace=False))
    return out


def _dedupe(inits):
    """Index of the first identical earlier restart for each restart (bit-identical runs)."""
    seen, first = {}, []
    for i, idx in enumerate(inits):
        key = np.asarray(idx, dtype=np.int64).tobytes()
        first.append(seen.setdefault(key, i))
    return first


class KMeansLeida:
    """K-Means with cosine distance on the GPU; same surface as the reference class
    (clustering/_clustering.py:24-247): KMeansLeida(k, metric, n_init, n_iter), fit(y, random_state),
    predict, transform, cluster_centers_, labels_, inertia_.

    Extra keyword arguments of fit(): init_indices (list of n_init index arrays, bypasses the
    legacy-RNG draw), centroid_update ('exact' | 'reference'), engine (reuse device
    data).  metric='cosine' (the LEiDA path) runs on the tensor cores; metric='euclidean' (accepted by the reference,
    :61-75, never used by its pipeline) runs the CUDA-core assignment with the same fp64 re-check discipline.
    """

    def __init__(self, k=2, metric="cosine", n_init=10, n_iter=1_000):
        if not isinstance(metric, str):
            raise TypeError("'metric' must be a string!")
        if metric not in ["euclidean", "cosine"]:
            raise ValueError("'metric' must be either 'cosine' or 'euclidean'")
        for name, val in {"k": k, "n_init": n_init, "n_iter": n_iter}.items():
            if not isinstance(val, int):
                raise TypeError(f"'{name}' must be an integer")
        if k < 2:
            raise ValueError("'k' must be > 1")
        self._k_, self._metric_, self._n_iter_, self._n_init_ = k, metric, n_iter, n_init

    def fit(self, y, random_state=None, init_indices=None, centroid_update="exact", engine=None):
        eng = engine if engine is not None else KMeansEngine(y)
        M = eng.m_global
        if init_indices is not None:
            inits = list(init_indices)
        elif eng.comm.world > 1:                       # one draw for the whole job (rank 0's RNG), see sweep_runs
            flat = np.zeros(self._n_init_ * self._k_, dtype=np.int64)
            if eng.comm.rank == 0:
                flat[:] = np.concatenate(_draw_inits(M, self._k_, self._n_init_, random_state))
            flat = eng.comm.bcast0_int64(flat)
            inits = [flat[r * self._k_:(r + 1) * self._k_].copy() for r in range(self._n_init_)]
        else:
            inits = _draw_inits(M, self._k_, self._n_init_, random_state)
        if len(inits) != self._n_init_:
            raise ValueError("init_indices must hold n_init index arrays")
        first = _dedupe(inits)
        uniq = sorted(set(first))
        res = eng.run([(self._k_, inits[i]) for i in uniq], n_iter=self._n_iter_, centroid_update=centroid_update,
                      metric=self._metric_)
        fitted = select_models(eng, res, {self._k_: [uniq.index(f) for f in first]}, self._n_init_)[self._k_]
        self.__dict__.update({k: v for k, v in fitted.__dict__.items() if k not in ("_k_", "_metric_", "_n_iter_", "_n_init_")})
        return self

    def _finish(self, eng, res, r, best_init, inertias, order, labels_device, centers_device, to_host=True):
        self.labels_device_ = labels_device
        self.centers_device_ = centers_device
        self._centers_exact = lambda: res.centers_exact(r)[order, :]
        if to_host:
            self.labels_ = labels_device.cpu().numpy().astype(np.int64)
            self.cluster_centers_ = centers_device.cpu().numpy()
        self.inertia_ = np.float64(inertias[best_init])
        self.n_passes_ = res.passes(r)
        self.inertia_runs_ = inertias
        self.best_init_ = best_init
        self._is_fitted = True

    def __getstate__(self):
        """Pickle like the reference's models (clustering/_clustering.py:345-350): host attributes only."""
        drop = ("labels_device_", "centers_device_", "_centers_exact")
        return {k: v for k, v in self.__dict__.items() if k not in drop}

    def _check_is_fitted(self):
        if not hasattr(self, "_is_fitted"):
            raise Exception("You have to fit the model first by using the 'fit' method.")

    def transform(self, y, closest=False):
        """Distances to the centroids in the model's metric, fp64 (clustering/_clustering.py:162-187)."""
        self._check_is_fitted()
        torch = nv.require_cuda()
        eng = y if isinstance(y, KMeansEngine) else KMeansEngine(y, n_features=self.cluster_centers_.shape[1])
        c = torch.from_numpy(np.ascontiguousarray(self.cluster_centers_, dtype=np.float32)).to(eng.X.device)
        euclid = self._metric_ == "euclidean"
        if euclid and eng.scale != 1.0:
            c = c * (1.0 / eng.scale)
        d = torch.empty((eng.M, c.shape[0]), dtype=torch.float64, device=eng.X.device)
        nv.call("leida_kmeans_transform_f64", nv.ptr(eng.X), eng.M, eng.N, eng.XP, nv.ptr(c), c.shape[0], c.shape[1],
                nv.ptr(d), None, 1 if euclid else 0, nv.stream())
        d = d.cpu().numpy()
        if euclid and eng.scale != 1.0:
            d = d * eng.scale
        return d.min(1) if closest else d

    def predict(self, y):
        """Closest centroid of each row (clustering/_clustering.py:189-207)."""
        self._check_is_fitted()
        eng = y if isinstance(y, KMeansEngine) else KMeansEngine(y, n_features=self.cluster_centers_.shape[1])
        return eng.predict(self.cluster_centers_, metric=self._metric_).cpu().numpy().astype(np.int64)


# pickles of fitted models name the reference's module path (see clustering/_clustering.py in this package)
KMeansLeida.__module__ = "pyleida.clustering._clustering"


def score_sample(M, drawn=None, comm=None):
    """Rows scored by Dunn / silhouette: all of them up to 30 000, else the reference's 30 000-row random
    sub-sample (clustering/_clustering.py:318-320) -- `drawn` when the draw was already made.  With several ranks the
    draw is rank 0's (every process has its own global RNG)."""
    if M <= 30_000:
        # the reference's two tests (N_samples>30_000 to draw, N_samples<30_000 to use everything) leave
        # M == 30 000 undefined (NameError); everything is scored here
        return None
    if drawn is not None:
        return drawn
    multi = comm is not None and comm.world > 1
    idx = np.zeros(30_000, dtype=np.int64)
    if not multi or comm.rank == 0:
        idx[:] = np.random.choice(np.arange(M), size=30_000, replace=False)
    return comm.bcast0_int64(idx) if multi else idx


def sweep_runs(M, K_min, K_max, n_init, random_state=None, init_indices=None, comm=None):
    """The (K, initial rows) runs of a K sweep in the reference's draw order (clustering/_clustering.py:
    318-329), identical restarts removed.  Returns (runs, plan): plan[k][r] = index into runs of
    restart r of K=k.  With several ranks the draws are made on rank 0 only and broadcast: every process has
    its own numpy global RNG, and ranks that seeded different initial rows would reduce unrelated centroids."""
    sweep_runs.last_sample = None
    multi = comm is not None and comm.world > 1
    drawn = None
    if init_indices is None and (not multi or comm.rank == 0):
        drawn = {}
        if M > 30_000:
            # :318-320 -- the reference draws its score sub-sample here; repeating the draw keeps the
            # global RNG stream (and therefore every later initial centroid) aligned with it.
            sweep_runs.last_sample = np.random.choice(np.arange(M), size=30_000, replace=False)
        for k in range(K_min, K_max + 1):
            drawn[k] = _draw_inits(M, k, n_init, random_state)
    if init_indices is None and multi:
        ks = list(range(K_min, K_max + 1))
        n_sample = 30_000 if M > 30_000 else 0
        flat = np.zeros(n_sample + n_init * sum(ks), dtype=np.int64)
        if comm.rank == 0:
            parts = [sweep_runs.last_sample] if n_sample else []
            parts += [np.asarray(idx, dtype=np.int64) for k in ks for idx in drawn[k]]
            flat[:] = np.concatenate(parts)
        flat = comm.bcast0_int64(flat)
        sweep_runs.last_sample = flat[:n_sample].copy() if n_sample else None
        drawn, o = {}, n_sample
        for k in ks:
            drawn[k] = [flat[o + r * k: o + (r + 1) * k].copy() for r in range(n_init)]
            o += n_init * k
    runs, plan = [], {}
    for k in range(K_min, K_max + 1):                                          # :326
        inits = list(init_indices[k]) if init_indices is not None else drawn[k]
        if len(inits) != n_init:
            raise ValueError(f"init_indices[{k}] must hold n_init index arrays")
        first = _dedupe(inits)
        uniq = sorted(set(first))
        base = len(runs)
        runs.extend((k, inits[i]) for i in uniq)
        plan[k] = [base + uniq.index(f) for f in first]
    return runs, plan


def select_models(engine, res, plan, n_init, to_host=True, stacked=False):
    """Best restart per K (:139-146) + relabelling (:234-247) -> {k: fitted KMeansLeida}.  The choices are host
    decisions on a few numbers per run; the device work -- relabelling the chosen run of every K and picking its
    centroid rows -- is one upload and one launch for all K.  stacked=True also returns the (n_K, M) int32 device
    array of all relabelled label rows (K in plan order)."""
    torch = engine.torch
    ks = list(plan.keys())
    picks = []
    for k in ks:
        inertias = [res.inertia(r) for r in plan[k]]
        best_init = _select_best(inertias)
        r = plan[k][best_init]
        lut, order = _remap_permutation(res.counts(r))
        picks.append((k, r, best_init, inertias, lut, order))
    kmax = max(len(p[4]) for p in picks)
    n_out = len(picks)
    row_of = np.asarray([res._inv[p[1]] for p in picks], dtype=np.int32)
    luts = np.zeros((n_out, kmax), dtype=np.int32)
    crow = []
    for c, p in enumerate(picks):
        luts[c, :len(p[4])] = p[4]
        j = res._inv[p[1]]
        crow.append(int(res._crow0[j]) + np.asarray(p[5], dtype=np.int64))
    # one pinned-free upload: [src rows | luts | centroid rows]
    blob = np.concatenate([row_of.astype(np.int64), luts.reshape(-1).astype(np.int64), np.concatenate(crow)])
    d_blob = torch.from_numpy(blob).to(res._labels.device)
    d_rows = d_blob[:n_out].to(torch.int32)
    d_luts = d_blob[n_out:n_out + n_out * kmax].to(torch.int32)
    d_crow = d_blob[n_out + n_out * kmax:]
    M = int(res._labels.shape[1])
    labels_all = torch.empty((n_out, M), dtype=torch.int32, device=res._labels.device)
    nv.call("leida_remap_labels_batch_i32", nv.ptr(res._labels), M, n_out, nv.ptr(d_rows), nv.ptr(d_luts), kmax,
            nv.ptr(labels_all), nv.stream())
    centers_all = res._cent[d_crow][:, :engine.N]                   # (sum K, N) in the relabelled order
    if engine.scale != 1.0:
        centers_all = centers_all * engine.scale
    models, off = {}, 0
    for c, (k, r, best_init, inertias, lut, order) in enumerate(picks):
        km = KMeansLeida(k=k, n_init=n_init, n_iter=1_000)
        km._finish(engine, res, r, best_init, inertias, order, labels_all[c], centers_all[off:off + len(order)], to_host)
        off += len(order)
        models[k] = km
    return (models, labels_all) if stacked else models


def identify_states(eigens_dataset, K_min=2, K_max=20, n_init=15, random_state=None, plot=True, save_results=False,
                    path=None, init_indices=None, centroid_update="exact", engine=None, scores=True):
    """K sweep of clustering/_clustering.py:250-374 with every (K, restart) run advanced in
    lock-step on the GPU.  Returns (predictions, clustering_performance, models) like the reference.

    eigens_dataset: DataFrame with 'subject_id', 'condition' and one column per ROI (the reference's
    input), or a (metadata DataFrame, KMeansEngine) pair through `engine` to reuse device data.
    scores=True computes the Dunn / silhouette / Davies-Bouldin columns (:331-339) on the GPU
    (clustering/scores.py; row-sharded engines gather the scored sample); `plot` is accepted and ignored.
    init_indices: optional {k: [idx_r ...]} replacing the legacy-RNG draws.
    """
    import pandas as pd
    if save_results and path is None:
        raise ValueError("You must provide a path to save the results")
    meta = eigens_dataset[["subject_id", "condition"]].copy()
    if engine is None:
        X = np.array(eigens_dataset.iloc[:, 2:], dtype=np.float32)            # :316
        engine = KMeansEngine(X)
    runs, plan = sweep_runs(engine.m_global, K_min, K_max, n_init, random_state, init_indices, comm=engine.comm)
    res = engine.run(runs, n_iter=1_000, centroid_update=centroid_update)
    by_k = select_models(engine, res, plan, n_init)
    ks = list(range(K_min, K_max + 1))
    sc = None
    if scores:
        from .scores import clustering_scores
        torch = engine.torch
        sc = clustering_scores(engine, torch.stack([by_k[k].labels_device_ for k in ks]), ks,
                               [by_k[k]._centers_exact() for k in ks],
                               score_sample(engine.m_global, sweep_runs.last_sample, engine.comm))
    models, perf = {}, []
    for c, k in enumerate(ks):
        km = by_k[k]
        models[f"k_{k}"] = km
        perf.append({"k": k, "Dunn_score": sc["Dunn_score"][c] if sc else np.nan, "distortion": km.inertia_,
                     "silhouette_score": sc["silhouette_score"][c] if sc else np.nan,
                     "Davies-Bouldin_score": sc["Davies-Bouldin_score"][c] if sc else np.nan})
        meta[f"k_{k}"] = km.labels_                                            # :342
    identify_states.last_batch = res
    perf = pd.DataFrame(perf)
    if save_results:
        from ..results_io import save_clustering
        save_clustering(path, meta, perf, models)
    return meta, perf, models


def patterns_stability(X, y=None, n_clusters=None, folds=5, metric="ari", plot=True, darkstyle=False):
    """Stability of the cluster assignments across folds (clustering/_clustering.py:471-605): half of the
    rows are held out, a KMeansLeida model is fitted on every training fold of the other half (on the GPU),
    each model labels the held-out rows, and the labelings are compared pairwise with ARI or AMI.
    The random split and the k-means restarts consume numpy's global RNG in the reference's order, so
    np.random.seed(s) reproduces the reference's result.  `plot` / `darkstyle` are accepted and ignored.
    Returns the (folds, folds) score matrix."""
    from sklearn.metrics import adjusted_mutual_info_score, adjusted_rand_score
    from sklearn.model_selection import KFold, StratifiedKFold, train_test_split
    if not isinstance(metric, str) or metric not in ["ari", "ami"]:
        raise Exception("'metric' must be either 'ari' or 'ami'")
    if not isinstance(X, np.ndarray):
        raise TypeError("'X' must be a 2D array!")
    if y is None and n_clusters is None:
        raise Exception("You must specify either predicted labels in 'y' or a number of clusters ('n_clusters').")
    if y is not None:
        if not isinstance(y, np.ndarray):
            raise TypeError("'y' must be an array!")
        n_clusters = int(np.unique(y).size)
        X_tr, X_ts, y_tr, _ = train_test_split(X, y, stratify=y, train_size=.5)           # :544
        splits = StratifiedKFold(n_splits=folds).split(X_tr, y_tr)
    else:
        X_tr, X_ts = train_test_split(X, train_size=.5)                                    # :546
        splits = KFold(n_splits=folds).split(X_tr)
    test_engine = KMeansEngine(np.ascontiguousarray(X_ts, dtype=np.float32))
    predictions = np.empty((X_ts.shape[0], folds))
    for i, (tr_idx, _) in enumerate(splits):                                              # :551-562
        km = KMeansLeida(k=n_clusters)
        km.fit(np.ascontiguousarray(X_tr[tr_idx], dtype=np.float32))
        predictions[:, i] = km.predict(test_engine)
    score = adjusted_rand_score if metric == "ari" else adjusted_mutual_info_score
    scores = np.empty((folds, folds))
    for a in range(folds):                                                                # :566-578
        for b in range(folds):
            scores[a, b] = score(predictions[:, a], predictions[:, b])
    return scores
