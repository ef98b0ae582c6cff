"""Short program for ncu: one launch of the persistent GA kernel (N generations) at config3 size."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import allhic_b200 as A  # noqa: E402
from allhic_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config3"
gens = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cfg = synth.CONFIGS[name]
ids, clmf, _ = synth.materialize(name)
clm = A.CLM(clmf, ids)
eng = A.Engine(0)
eng.load(clm)
n, P = cfg["n"], cfg["npop"]
if "--npop" in sys.argv:
    P = int(sys.argv[sys.argv.index("--npop") + 1])
rng = np.random.default_rng(1000)
init = rng.permutation(n).astype(np.int32)
eng.ga_begin(init, eng.ga_params(npop=P, ngen=1 << 30, max_gens=1 << 40, seed=42))
eng.ga_advance(gens)          # warm-up launch
import time
t0 = time.perf_counter()
st = eng.ga_advance(gens)     # profiled launch
dt = time.perf_counter() - t0
print("GA %s: %d generations in %.3f ms -> %.1f us/generation, device_ms total %.3f" % (name, gens, dt * 1e3, dt / gens * 1e6, st.device_ms))
eng.ga_end()
eng.close()
