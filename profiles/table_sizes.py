"""Entries of the resident kernel's pair table per tours-per-batch T (stored pairs + padding)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import allhic_b200 as A  # noqa: E402
from allhic_b200 import synth  # noqa: E402

for name in sys.argv[1:] or ["config3"]:
    ids, clmf, _ = synth.materialize(name)
    clm = A.CLM(clmf, ids)
    eng = A.Engine(0)
    eng.load(clm)
    print(name, "N", clm.N, "E", len(clm.tables()["pa"]), {t: eng.pair_table_entries(t) for t in (1, 2, 4)}, "kernel", eng.eval_m_kernel(synth.CONFIGS[name]["npop"]))
