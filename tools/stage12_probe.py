"""Stage 1+2 (Hilbert -> (cos, sin) -> rank-2 eigenvector) timing on config-3-shaped data, for a few kernel shapes.
    python tools/stage12_probe.py [subjects]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
from pyleida import synthetic as syn
from pyleida.signal_tools.phase import leading_eigenvectors, hilbert_phase
S = int(sys.argv[1]) if len(sys.argv) > 1 else 250
for N, T in ((400, 1200), (116, 1200)):
    ts, _ = syn.make_shard(S, N, T, 0, S, "states")
    x = torch.from_numpy(np.stack(list(ts.values()))).cuda()
    M = S * (T - 2)
    xp = (N + 31) // 32 * 32
    alg = 8 * S * N * T + 8 * M * N + 4 * M * xp
    lei = torch.empty((M, N), dtype=torch.float64, device="cuda")
    xf = torch.empty((M, xp), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for batch, threads in ((None, None), (1, 96), (1, 128), (2, 160), (2, 256), (3, 256), (4, 256)):
        for k, v in (("LEIDA_HILBERT_BATCH", batch), ("LEIDA_HILBERT_THREADS", threads)):
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = str(v)
        ts_ms = []
        for rep in range(4):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); leading_eigenvectors(x, out_f64=lei, out_f32=xf); e1.record(); torch.cuda.synchronize()
            ts_ms.append(e0.elapsed_time(e1))
        t = min(ts_ms[1:])
        print(f"N={N} S={S} batch={batch} threads={threads}: {t:.3f} ms  {alg/t/1e6:.0f} GB/s algorithmic "
              f"({alg/t/1e6/6537.6:.2f} of HBM peak); full config-3-size estimate {t*1000/S:.2f} ms" if N == 400 else
              f"N={N} S={S} batch={batch} threads={threads}: {t:.3f} ms  {alg/t/1e6:.0f} GB/s algorithmic ({alg/t/1e6/6537.6:.2f} of HBM peak)", flush=True)
    os.environ.pop("LEIDA_HILBERT_BATCH", None); os.environ.pop("LEIDA_HILBERT_THREADS", None)
    # correctness against the oracle on two subjects
    import leida_oracle as O
    got = lei[: 2 * (T - 2)].cpu().numpy()
    ref = np.vstack([O.leading_eigvec_rank2(O.hilbert_phase(np.asarray(v))) for v in list(ts.values())[:2]])
    print(f"  max |v - v_oracle| = {np.abs(got - ref).max():.2e}")
    del x, lei, xf
