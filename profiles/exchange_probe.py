"""Cost of one island exchange on a single GPU: ahx_ga_get_elites(2) + ahx_ga_put_migrants(14) (what rank r does at N = 8)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import allhic_b200 as A
from allhic_b200 import synth
ids, clmf, _ = synth.materialize("config3")
clm = A.CLM(clmf, ids); eng = A.Engine(0); eng.load(clm)
n, P = 2000, 8000
rng = np.random.default_rng(0)
eng.ga_begin(rng.permutation(n).astype(np.int32), eng.ga_params(npop=P, ngen=1 << 30, max_gens=1 << 40, seed=1))
eng.ga_advance(50)
mig = np.stack([rng.permutation(n) for _ in range(14)]).astype(np.int32)
for _ in range(3):
    eng.ga_get_elites(2); eng.ga_put_migrants(mig)
t0 = time.perf_counter()
for _ in range(20):
    eng.ga_get_elites(2)
t1 = time.perf_counter()
for _ in range(20):
    eng.ga_put_migrants(mig)
t2 = time.perf_counter()
print("get_elites(2): %.3f ms   put_migrants(14): %.3f ms" % ((t1 - t0) / 20 * 1e3, (t2 - t1) / 20 * 1e3))
assert eng.ga_selfcheck() == 0
eng.ga_end(); eng.close()
