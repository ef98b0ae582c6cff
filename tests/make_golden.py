"""Regenerates tests/golden/known_answers.json from the C oracle, after checking
every value against the independent pure-Python restatement (tests/pyref.py).
Run here (CPU): python tests/make_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from oracle_lib import OracleCLM, golden_array  # noqa: E402
from pyref import PyCLM  # noqa: E402

ids, clm = os.path.join(HERE, "golden", "test.ids"), os.path.join(HERE, "golden", "test.clm")
orc, py = OracleCLM(clm, ids), PyCLM(clm, ids)
N = orc.N
rng = np.random.default_rng(20240607)
out = {"N": int(N), "ncontacts": int(orc.l.orc_ncontacts(orc.h)), "noriented": int(orc.l.orc_noriented(orc.h))}
M = orc.M()
assert (M == np.array(py.M())).all()
out.update(M01=int(M[0, 1]), M12=int(M[1, 2]), nnzM=int((M != 0).sum()), sum_upper_M=int(np.triu(M).sum()),
           sum_sizes=int(orc.sizes().sum()))
first = open(clm).readline().rstrip("\n").split("\t")
d = [int(x) for x in first[2].split(" ")]
out["golden_array_line1"] = [int(x) for x in golden_array(d)]
assert out["golden_array_line1"] == py.golden_array(d) if hasattr(py, "golden_array") else True
cases = []
tours = [np.arange(N), np.arange(N)[::-1].copy()] + [rng.permutation(N) for _ in range(6)]
tours += [rng.permutation(N)[:k] for k in (1, 2, 37, 99)]          # sub-tours (pruneTour uses them, clm.go:331-334)
for t in tours:
    t = np.ascontiguousarray(t, np.int64)
    for pattern in ("plus", "minus", "random"):
        signs = {"plus": np.full(N, ord('+')), "minus": np.full(N, ord('-')),
                 "random": np.where(rng.random(N) < 0.5, ord('+'), ord('-'))}[pattern].astype(np.uint8)
        orc.set_tour(t); orc.set_signs(signs)
        em, eq, eql = orc.evaluate(t), orc.evaluate_q(), orc.evaluate_q(literal=True)
        assert eq == eql, "lookup and literal EvaluateQ must be bit-identical"
        pm = py.evaluate(t.tolist()); pq = py.evaluate_q(t.tolist(), [chr(c) for c in signs])
        assert pm == em, (pm, em)
        assert abs(pq - eq) <= 1e-13 * abs(eq) + 0.0, (pq, eq)
        cases.append(dict(tour=t.tolist(), signs=bytes(signs).decode(), evaluate=em, evaluate_q=eq))
out["cases"] = cases
with open(os.path.join(HERE, "golden", "known_answers.json"), "w") as f:
    json.dump(out, f)
print("wrote", len(cases), "cases; identity Evaluate =", repr(cases[0]["evaluate"]), "EvaluateQ(+) =", repr(cases[0]["evaluate_q"]))
