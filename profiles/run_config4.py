"""BASELINE configs[3]: 48 synthetic linkage groups x ~800 contigs farmed one group per GPU.
  python profiles/run_config4.py [n_groups] [--cpu K] [--per-group-process]   (uses every visible GPU)
Prints one JSON line: groups/hour, per-group seconds and final scores, and -- for K groups --
the CPU restatement's GA score on the same input (seed 42, reference defaults)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from allhic_b200 import farm, synth, _ffi  # noqa: E402

n_groups = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 48
n_cpu = int(sys.argv[sys.argv.index("--cpu") + 1]) if "--cpu" in sys.argv else 2
cfgs = synth.config4_groups()[:n_groups]
outdir = os.path.join(os.environ.get("TMPDIR", "/tmp"), "allhic_b200_config4")
os.makedirs(outdir, exist_ok=True)
def _make(c):
    ids, clm, _ = synth.materialize(c["n"], seed=c["seed"], outdir=outdir)
    return ids, clm


from concurrent.futures import ProcessPoolExecutor  # noqa: E402
with ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:   # the generator is single-threaded numpy
    groups = list(ex.map(_make, cfgs))
ndev = _ffi.lib().ahx_device_count()
t0 = time.perf_counter()
persistent = "--per-group-process" not in sys.argv
res = farm.optimize_groups(groups, list(range(max(1, ndev))), persistent=persistent)
wall = time.perf_counter() - t0
out = {"workload": "config4: %d synthetic linkage groups, N ~ U[700,900], full `optimize` (2 GA phases, npop 100, ngen 5000, flips)" % n_groups,
       "n_gpus": ndev, "processes": "one optimize-many worker per GPU" if persistent else "one optimize process per group", "wall_s": wall, "groups_per_hour": 3600.0 * n_groups / wall, "all_ok": all(r["ok"] for r in res),
       "seconds_per_group": [round(r["seconds"], 2) for r in res], "final_scores": [r["final_score"] for r in res],
       "n_contigs": [r["n_contigs"] for r in res]}
if n_cpu:
    import numpy as np
    from oracle_lib import OracleCLM
    cpu = []
    for g in range(min(n_cpu, n_groups)):
        orc = OracleCLM(groups[g][1], groups[g][0])
        orc.activate(shuffle=True, seed=42, flip_all=False)
        t1 = time.perf_counter()
        r = orc.ga_run(npop=100, ngen=5000, mutprob=0.2, seed=42, max_gens=1000000, max_seconds=60.0, nthreads=os.cpu_count() or 1)
        cpu.append({"group": g, "cpu_ga_phase1_score": r.best_score, "cpu_ga_phase1_seconds": time.perf_counter() - t1, "generations": r.generations})
    out["cpu_restatement"] = cpu
print(json.dumps(out))
