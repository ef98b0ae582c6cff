"""Small end-to-end exercise (evaluate resident / streamed / sub-tour, persistent GA both ways, flips): the target for
compute-sanitizer where that tool is available (it is closed on the build pool), and a quick consistency check otherwise."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import allhic_b200 as A
gold = os.path.join(ROOT, "tests", "golden")
clm = A.CLM(os.path.join(gold, "test.clm"), os.path.join(gold, "test.ids"))
eng = A.Engine(0); eng.load(clm)
rng = np.random.default_rng(0)
n = clm.N
tours = np.stack([rng.permutation(n) for _ in range(37)]).astype(np.int32)
a = eng.eval_m(tours)
os.environ["AHX_EVAL_STREAM"] = "1"
b = eng.eval_m(tours)
os.environ["AHX_EVAL_STREAM"] = "0"
sub = np.stack([rng.permutation(n)[:40] for _ in range(9)]).astype(np.int32)
c = eng.eval_m(sub)
assert np.allclose(a, b, rtol=1e-13)
best, st = eng.ga_run(tours[0], eng.ga_params(npop=64, ngen=1 << 30, max_gens=40, seed=3))
os.environ["AHX_EVAL_STREAM"] = "1"
best2, st2 = eng.ga_run(tours[0], eng.ga_params(npop=64, ngen=1 << 30, max_gens=40, seed=3))
os.environ["AHX_EVAL_STREAM"] = "0"
assert st.best_score == st2.best_score or abs(st.best_score - st2.best_score) < 1e-12
signs = np.where(rng.random((3, n)) < 0.5, ord('+'), ord('-')).astype(np.uint8)
q = eng.eval_q(tours[0], signs)
s2, *_ = eng.flip_whole(best, signs[0]); eng.flip_one(best, s2)
print("sanitizer target ok", st.best_score, float(a[0]), float(c[0]), float(q[0]))
eng.close()
