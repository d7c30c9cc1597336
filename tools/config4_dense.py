"""BASELINE config 4 (200 subjects x 1000 ROIs x 1200 volumes) through the reference's LITERAL stage 2 -- materialise
dFC (N, N, T-2) per subject (phase_coherence, _signal_tools.py:137-183) and run the general symmetric eigensolver on it
(get_eigenvectors, :185-224) -- next to the rank-2 closed form the pipeline uses.  dFC is 9.58 GB per subject, so the
subjects go through one at a time; a few subjects are timed and the total extrapolated.
    python tools/config4_dense.py [n_subjects_timed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leida-python_b200"))
import numpy as np, torch
from pyleida import synthetic as syn
from pyleida.signal_tools.phase import hilbert_phase, phase_coherence, get_eigenvectors, leading_eigenvectors
n_sub = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N, T, S = 1000, 1200, 200
ts, _ = syn.make_shard(S, N, T, 0, n_sub, "states")
xs = [torch.from_numpy(v).cuda() for v in ts.values()]
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
tot = {"hilbert": 0.0, "dfc": 0.0, "eig": 0.0, "rank2": 0.0}
for i, x in enumerate(xs):
    e0 = ev(); ph = hilbert_phase(x); e1 = ev(); dfc = phase_coherence(ph); e2 = ev()
    lei, info = get_eigenvectors(dfc, return_info=True); e3 = ev()
    r2, _ = leading_eigenvectors(x.unsqueeze(0), want_f32=False); e4 = ev()
    torch.cuda.synchronize()
    err = float((lei - r2).abs().max())
    t = [e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e3.elapsed_time(e4)]
    print(f"subject {i}: hilbert {t[0]:.2f} ms, dFC write {t[1]:.2f} ms ({dfc.numel()*8/1e9:.2f} GB -> {dfc.numel()*8/t[1]/1e6:.0f} GB/s), "
          f"dense eigensolver {t[2]:.1f} ms ({info['iterations']} iterations, {info['unconverged']} unconverged; "
          f"{info['iterations']*dfc.numel()*8/t[2]/1e6:.0f} GB/s of dFC reads), rank-2 path {t[3]:.2f} ms; max |dense - rank2| = {err:.2e}", flush=True)
    if i > 0:
        for k, v in zip(tot, t): tot[k] += v
    del dfc, lei
n = max(1, n_sub - 1)
print(f"per subject (first one excluded): " + ", ".join(f"{k} {v/n:.2f} ms" for k, v in tot.items()))
print(f"config 4, 200 subjects, stage 1+2: literal path {(tot['hilbert']+tot['dfc']+tot['eig'])/n*200/1e3:.2f} s (EXTRAPOLATED from {n} subjects), "
      f"rank-2 path {tot['rank2']/n*200/1e3:.3f} s")
