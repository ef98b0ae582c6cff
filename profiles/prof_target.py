"""Short program for `ncu --set full`: a few launches of each hot kernel at
config3 size (N=2000, P=8000), nothing else.

  python profiles/prof_target.py [config3] [converged]
    converged: tours are small perturbations of the identity (what the GA holds
    once it has converged: nearly every stored pair is inside the LIMIT window)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import allhic_b200 as A  # noqa: E402
from allhic_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config3"
converged = "converged" in sys.argv[2:]
cfg = synth.CONFIGS[name]
ids, clmf, _ = synth.materialize(name)
clm = A.CLM(clmf, ids)
eng = A.Engine(0)
eng.load(clm)
n, P = cfg["n"], cfg["npop"]
rng = np.random.default_rng(1000)
if converged:
    tours = np.repeat(np.arange(n, dtype=np.int32)[None, :], P, 0)
    for t in tours:                                   # a handful of random swaps each
        i, j = rng.integers(0, n, (2, 8))
        t[i], t[j] = t[j], t[i]
else:
    tours = np.stack([rng.permutation(n) for _ in range(P)]).astype(np.int32)
eng.pop_set(tours)
for _ in range(3):
    eng.flush_l2()
    ms = eng.pop_eval_m("sparse", 1)
print("eval_m %s: %.3f ms -> %.2f M evals/s" % ("converged" if converged else "random", ms, P / ms / 1e3))
eng.ga_begin(tours[0], eng.ga_params(npop=P, ngen=1 << 30, max_gens=1 << 40, seed=42))
st = eng.ga_advance(2)
eng.ga_end()
signs = np.where(rng.random((256, n)) < 0.5, ord('+'), ord('-')).astype(np.uint8)
eng.eval_q(tours[0], signs)
s2, *_ = eng.flip_whole(tours[0], signs[0])
eng.flip_one(tours[0], s2)
print("launches", eng.launch_count())
eng.close()
