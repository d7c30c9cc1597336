"""Measure the error of the tensor-core (bf16 x 3, fp32 accumulate in TMEM) score against fp64.

For N in {116, 400, 1000} and clustered / iid rows: run two Lloyd passes of the K=2..20 x 15-restart sweep to
get realistic centroids, pack them like a production pass (leida_kmeans_tc_plan), dump the raw scores with
leida_kmeans_tc_scores and compare with  dot64(x, c) / |c|  (fp32 inputs, fp64 arithmetic).  The error is
reported in units of |x| -- the unit of the trust threshold tau (kmeans_tc.cu: default_tau).

    python tools/tc_score_error.py [rows_per_case]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pyleida import _native as nv  # noqa: E402
from pyleida import synthetic as syn  # noqa: E402
from pyleida.clustering.kmeans import KMeansEngine, sweep_runs  # noqa: E402
from pyleida.signal_tools.phase import leading_eigenvectors  # noqa: E402


def score_errors(eng, cent, ks, crow0):
    """max / quantiles of |S_tc - S_64| / |x| over all (row, centroid) pairs of the packed runs."""
    dev = eng.X.device
    R, CR, N, XP = len(ks), int(ks.sum()), eng.N, eng.XP
    x1, x2, npk = eng._tc_operands()
    run_k = torch.from_numpy(ks.astype(np.int32)).to(dev)
    run_c0 = torch.from_numpy(crow0.astype(np.int32)).to(dev)
    state = torch.zeros((R, nv.STATE_WORDS), dtype=torch.int64, device=dev)
    state[:, nv.ST_ACTIVE] = 1
    cnorm = cent[:, :N].double().norm(dim=1).contiguous()
    bn = int(nv.lib.leida_kmeans_tc_tile_columns())
    c_rows = (4 * CR + bn * 17 + bn - 1) // bn * bn
    c1 = torch.zeros((c_rows, npk), dtype=torch.bfloat16, device=dev)
    c2 = torch.zeros((c_rows, npk), dtype=torch.bfloat16, device=dev)
    plan = torch.zeros(2, dtype=torch.int32, device=dev)
    tile_first = torch.zeros(c_rows // 128, dtype=torch.int32, device=dev)
    tile_count = torch.zeros(c_rows // 128, dtype=torch.int32, device=dev)
    prun = torch.zeros(R, dtype=torch.int32, device=dev)
    pcol = torch.zeros(R, dtype=torch.int32, device=dev)
    run_pcol = torch.zeros(R, dtype=torch.int32, device=dev)
    nv.call("leida_kmeans_tc_plan", R, nv.ptr(run_k), nv.ptr(run_c0), nv.ptr(state), nv.ptr(cent), nv.ptr(cnorm), N, XP,
            nv.ptr(plan), nv.ptr(tile_first), nv.ptr(tile_count), nv.ptr(prun), nv.ptr(pcol), nv.ptr(run_pcol),
            nv.ptr(c1), nv.ptr(c2), 1, nv.stream())
    labels_new = torch.empty((R, eng.M), dtype=torch.int32, device=dev)
    cap = 1 << 16
    worklist = torch.zeros(8 * cap + 8, dtype=torch.int32, device=dev)
    scores = torch.full((eng.M, c_rows), float("nan"), dtype=torch.float32, device=dev)
    nv.call("leida_kmeans_tc_scores", nv.ptr(x1), nv.ptr(x2), nv.ptr(eng.xnorm), eng.M, N, nv.ptr(c1), nv.ptr(c2), c_rows,
            nv.ptr(plan), nv.ptr(tile_first), nv.ptr(tile_count), nv.ptr(prun), nv.ptr(labels_new), nv.ptr(worklist), cap,
            nv.ptr(scores), c_rows, nv.stream())
    torch.cuda.synchronize()
    pc = run_pcol.cpu().numpy()
    X64 = eng.X[:, :N].double()
    xn = X64.norm(dim=1)
    C64 = cent[:, :N].double() / cnorm[:, None]
    worst, errs = 0.0, []
    for r in range(R):
        col = int(pc[r]) & 0xFFFFFF
        k, c0 = int(ks[r]), int(crow0[r])
        ref = X64 @ C64[c0:c0 + k].T                               # (M, k) fp64
        got = scores[:, col:col + k].double()
        e = ((got - ref).abs() / xn[:, None]).reshape(-1)
        worst = max(worst, float(e.max()))
        errs.append(e[:: max(1, e.numel() // 20000)].cpu())       # thinned sample for the quantiles
    e = torch.cat(errs).numpy()
    return worst, float(np.sqrt((e ** 2).mean())), float(np.quantile(e, 0.9999)), int(eng.M) * CR


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 24000
    print(f"{'N':>5} {'NPK':>5} {'kind':>7} {'pairs':>10} {'max err/|x|':>12} {'rms':>10} {'p99.99':>10} {'tau(N)':>10} {'tau/2/max':>9}")
    for N in (116, 400, 1000):
        for kind in ("states", "iid"):
            T = 1200
            S = max(2, rows // (T - 2))
            ts, _ = syn.make_shard(S, N, T, 0, S, kind)
            x = torch.from_numpy(np.stack(list(ts.values()))).cuda()
            _, xf = leading_eigenvectors(x, want_f64=False)
            eng = KMeansEngine(xf, n_features=N)
            inits = syn.init_indices(eng.M, 2, 20, 15, seed=0)
            runs, _ = sweep_runs(eng.M, 2, 20, 15, None, inits)
            res = eng.run(runs, n_iter=2)                          # centroids after two Lloyd updates
            worst, rms, q, pairs = score_errors(eng, res._cent, res._ks.astype(np.int64), res._crow0.astype(np.int64))
            tau = float(nv.lib.leida_kmeans_tc_tau(N))
            npk = int(nv.lib.leida_kmeans_tc_pitch(N))
            print(f"{N:5d} {npk:5d} {kind:>7} {pairs:10d} {worst:12.3e} {rms:10.3e} {q:10.3e} {tau:10.3e} {tau / 2 / worst:9.1f}",
                  flush=True)
            del eng, res, x, xf


if __name__ == "__main__":
    main()
