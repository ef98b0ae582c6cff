"""GPU: `allhic-b200 optimize` as a drop-in for `allhic optimize`: the .tour record
sequence, the stdout progress lines, --skipGA, --resume with its quirks, and the hand-off of the last
record to `build`; `optimize-many` with several groups in flight and a shared genome-wide CLM parse;
`optimize --gpus N` (island model)."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from conftest import FIX_CLM, FIX_IDS
import oracle_lib as O

pytestmark = pytest.mark.gpu


def _binary():
    from allhic_b200 import farm
    return farm.BINARY


def _records(path):
    """[(label, [words])] of a tour file (optimize.go:176-184: `>LABEL\\n` + `name± name± ...\\n`)."""
    lines = open(path).read().split("\n")
    assert lines[-1] == ""
    lines = lines[:-1]
    assert len(lines) % 2 == 0
    recs = []
    for h, row in zip(lines[0::2], lines[1::2]):
        assert h.startswith(">") and " " not in h
        recs.append((h[1:], row.split(" ")))
    return recs


def _parse_last_tour(path):
    """What `build` reads (build.go:130-155 parseLastTour over optimize.go:98-117 parseTourFile): the last
    non-header row, split on single spaces, each word = name + strand in {+,-,?}."""
    words = None
    for row in open(path).read().split("\n"):
        row = row.strip()
        if row == "":
            continue
        if row[0] == ">":
            continue
        words = row.split(" ")
    out = []
    for w in words:
        at, ao = w[:-1], w[-1]
        out.append((at, ao) if ao in "+-?" else (w, "?"))
    return out


def _fixture_copy(d):
    d.mkdir(parents=True, exist_ok=True)
    shutil.copy(FIX_IDS, d / "test.ids"); shutil.copy(FIX_CLM, d / "test.clm")
    return str(d / "test.ids"), str(d / "test.clm")


def test_tour_file_record_sequence_and_stdout(built_lib, tmp_path):
    ids, clm = _fixture_copy(tmp_path / "a")
    r = subprocess.run([_binary(), "optimize", ids, clm, "--ngen", "3000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Success" in r.stderr
    orc = O.OracleCLM(clm, ids)
    names = orc.names()
    idx = {n: i for i, n in enumerate(names)}
    recs = _records(str(tmp_path / "a" / "test.tour"))
    labels = [l for l, _ in recs]
    # INIT, then per GA phase a record at generation 0 and at every 500th generation (evaluate.go:294-300),
    # then FLIPWHOLE{n} / FLIPONE{n} pairs until both REJECT (optimize.go:68-74); FINAL goes to stdout only
    assert labels[0] == "INIT"
    k = 1
    ga = {1: [], 2: []}
    while k < len(labels) and labels[k].startswith("GA"):
        m = re.fullmatch(r"GA([12])-(\d+)-(\d+\.\d{5})", labels[k])
        assert m, labels[k]
        ga[int(m.group(1))].append((int(m.group(2)), float(m.group(3)), recs[k][1]))
        k += 1
    for phase in (1, 2):
        gens = [g for g, _, _ in ga[phase]]
        assert gens == list(range(0, gens[-1] + 1, 500)) and len(gens) >= 2
        assert all(b[1] >= a[1] for a, b in zip(ga[phase], ga[phase][1:]))            # best-ever score never decreases
        for g, score, words in ga[phase]:
            t = np.array([idx[w[:-1]] for w in words], np.int64)
            assert sorted(t.tolist()) == list(range(len(names)))                       # bit-exact validity
            assert abs(-orc.evaluate(t) - score) <= 0.5000001e-5                        # the label is this tour's score, %.5f
    assert ga[2][0][1] >= ga[1][-1][1] - 1e-5                                           # phase 2 starts from phase 1's best tour
    rest = labels[k:]
    assert len(rest) % 2 == 0 and len(rest) >= 2
    assert rest == [x for n in range(1, len(rest) // 2 + 1) for x in ("FLIPWHOLE%d" % n, "FLIPONE%d" % n)]
    assert "FINAL" not in labels
    for _, words in recs:
        assert all(w[-1] in "+-" for w in words) and sorted(w[:-1] for w in words) == sorted(names)
    # flips change signs only: every record after the GA keeps the GA's order
    order = [w[:-1] for w in recs[k - 1][1]]
    assert all([w[:-1] for w in words] == order for _, words in recs[k:])
    # stderr: the reference's log lines; the last flip round is REJECT / REJECT
    log = r.stderr
    for needle in ("Parse REfile", "Parse clmfile", "Active tigs: 100 (length=9999998)", "GA initialized (npop: 100, ngen: 3000, mu: 0.20, rng: 42, break: 10000000)",
                   "FLIPALL", "FLIPWHOLE", "FLIPONE: N_accepts=", "Terminating ... no more ACCEPT", "Optimization history logged to"):
        assert needle in log, needle
    assert len(re.findall(r"FLIPONE \(\d+/100\)", log)) == 2 * (len(rest) // 2)        # every 50th step of every pass (orientation.go:103)
    # stdout: INIT record, one progress line per GA record, FINAL record
    out = r.stdout.split("\n")
    assert out[0] == ">INIT" and out[1] == " ".join(recs[0][1])
    prog = [l for l in out if l.startswith("Current iteration")]
    want = ["Current iteration GA%d-%d: max_score=%.5f" % (ph, g, s) for ph in (1, 2) for g, s, _ in ga[ph]]
    assert prog == want
    assert out[-3] == ">FINAL" and out[-2] == " ".join(recs[-1][1]) and out[-1] == ""
    # hand-off to `build`: the last record parses into (contig, strand) for every contig exactly once
    last = _parse_last_tour(str(tmp_path / "a" / "test.tour"))
    assert sorted(a for a, _ in last) == sorted(names) and all(s in "+-" for _, s in last)
    # the converged order is the simulated truth (or its mirror image) on this fixture
    t = [idx[a] for a, _ in last]
    assert t == list(range(100)) or t == list(range(99, -1, -1))
    # orientation: the greedy flips end where no single flip and no whole flip improves EvaluateQ
    final_signs = np.full(len(names), ord('+'), np.uint8)
    for a, s in last:
        final_signs[idx[a]] = ord(s)
    orc.set_tour(np.array(t, np.int64)); orc.set_signs(final_signs)
    q_final = orc.evaluate_q()
    for i in range(len(names)):
        flipped = final_signs.copy(); flipped[i] = ord('-') if flipped[i] == ord('+') else ord('+')
        orc.set_signs(flipped)
        assert orc.evaluate_q() <= q_final
    orc.set_signs(np.where(final_signs == ord('+'), ord('-'), ord('+')).astype(np.uint8))
    assert orc.evaluate_q() <= q_final


def test_skip_ga_and_determinism(built_lib, tmp_path):
    """--skipGA (functional-tests.sh:13): no GA records; the order stays the shuffled INIT order; same
    seed => byte-identical tour file, other seed => another shuffle."""
    runs = {}
    for tag, seed in (("a", "42"), ("b", "42"), ("c", "7")):
        ids, clm = _fixture_copy(tmp_path / tag)
        r = subprocess.run([_binary(), "optimize", ids, clm, "--skipGA", "--seed", seed], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "Success" in r.stderr and "GA initialized" not in r.stderr
        runs[tag] = open(str(tmp_path / tag / "test.tour")).read()
        recs = _records(str(tmp_path / tag / "test.tour"))
        assert [l for l, _ in recs][0] == "INIT" and not any(l.startswith("GA") for l, _ in recs)
        order = [w[:-1] for w in recs[0][1]]
        assert all([w[:-1] for w in words] == order for _, words in recs)
        assert "Current iteration" not in r.stdout
    assert runs["a"] == runs["b"] and runs["a"] != runs["c"]


def test_resume_backs_up_and_keeps_only_the_listed_contigs(built_lib, tmp_path):
    """--resume (optimize.go:44-51): the existing tour file is parsed (last record only), backed up to
    .tour.sav, and ONLY the contigs it lists stay active (clm.go:400-417: Activate starts from the active
    set -- in REfile order, reshuffled; the stored order is not kept, the stored signs are reset by
    flipAll's '+' initialisation).  The sub-tour GA runs on the persistent kernel."""
    ids, clm = _fixture_copy(tmp_path / "r")
    names = [l.split()[0] for l in open(ids) if not l.startswith("#")]
    keep = names[10:70]
    tour = tmp_path / "r" / "test.tour"
    old = ">OLD1\n%s\n>LAST\n%s ghost+\n" % (" ".join(n + "+" for n in names), " ".join(n + "-" for n in reversed(keep)))
    tour.write_text(old)
    r = subprocess.run([_binary(), "optimize", ids, clm, "--resume", "--ngen", "2000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Success" in r.stderr
    assert open(str(tour) + ".sav").read() == old
    assert "Found existing tour file" in r.stderr and "Backup" in r.stderr and "Active tigs: 60" in r.stderr
    assert "ERROR Contig ghost not found!" in r.stderr                                  # clm.go parseTourFile: unknown names are reported and skipped
    recs = _records(str(tour))
    assert recs[0][0] == "INIT"
    for _, words in recs:
        assert sorted(w[:-1] for w in words) == sorted(keep)
    out = r.stdout.split("\n")
    # stdout carries INIT twice: parseTourFile prints the parsed tour, Run prints the activated one
    assert [l for l in out if l.startswith(">")] == [">INIT", ">INIT", ">FINAL"]
    assert out[1] == " ".join(n + "-" for n in reversed(keep))
    last = [a for a, _ in _parse_last_tour(str(tour))]
    pos = [names.index(a) for a in last]
    assert pos == sorted(pos) or pos == sorted(pos, reverse=True)                       # converged on the sub-tour as well
    # without --resume an existing tour file is simply overwritten, no backup
    ids2, clm2 = _fixture_copy(tmp_path / "n")
    (tmp_path / "n" / "test.tour").write_text(old)
    r = subprocess.run([_binary(), "optimize", ids2, clm2, "--skipGA"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and not os.path.exists(str(tmp_path / "n" / "test.tour.sav"))
    assert "Active tigs: 100" in r.stderr


def test_cli_errors_are_fatal_like_the_reference(built_lib, tmp_path):
    ids, clm = _fixture_copy(tmp_path / "e")
    r = subprocess.run([_binary(), "optimize", ids, str(tmp_path / "missing.clm")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "CRITIC" in r.stderr and "Success" not in r.stderr
    r = subprocess.run([_binary(), "optimize", ids], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "accepts 2 arg(s), received 1" in r.stderr


def test_optimize_many_concurrent_groups_and_shared_parse(built_lib, tmp_path):
    """Several groups in flight on one GPU (--jobs), all naming ONE genome-wide clmfile (what `pipeline`
    hands every group, allhic.go:285-293): the clmfile is parsed once, and every group's tour file is
    byte-identical to a separate `optimize` run's."""
    from test_host import _write_genome
    ns = [60, 45, 80, 30]
    runs = {}
    for mode in ("many", "single"):
        d = tmp_path / mode
        d.mkdir()
        clmf, refiles = _write_genome(d, ns)
        if mode == "many":
            (d / "list").write_text("".join("%s %s\n" % (re_g, clmf) for re_g in refiles))
            r = subprocess.run([_binary(), "optimize-many", str(d / "list"), "--ngen", "300", "--jobs", "4"], capture_output=True, text=True, timeout=600)
            assert r.returncode == 0 and r.stderr.count("Success") == 4
            assert len(re.findall(r"Parsed clmfile `.*` once for 4 groups", r.stderr)) == 1
            done = re.findall(r"^AHX_GROUP_DONE (\d+) (\S+) device=0 seconds=[0-9.]+$", r.stdout, re.M)
            assert sorted(int(i) for i, _ in done) == [0, 1, 2, 3]
            # every group's stdout block is contiguous and complete
            blocks = re.split(r"^AHX_GROUP_DONE.*$", r.stdout, flags=re.M)
            assert all(b.count(">INIT") == 1 and b.count(">FINAL") == 1 for b in blocks[:4])
            # one bad group fails alone
            (d / "list2").write_text("%s %s\n%s %s\n" % (refiles[0], clmf, str(d / "nope.txt"), clmf))
            r2 = subprocess.run([_binary(), "optimize-many", str(d / "list2"), "--skipGA", "--jobs", "2"], capture_output=True, text=True, timeout=600)
            assert r2.returncode == 1 and "AHX_GROUP_DONE 0" in r2.stdout and "AHX_GROUP_FAILED 1" in r2.stdout
            subprocess.run([_binary(), "optimize", refiles[0], clmf, "--ngen", "300"], capture_output=True, text=True, timeout=600)   # restore group 0's tour
        else:
            for re_g in refiles:
                r = subprocess.run([_binary(), "optimize", re_g, clmf, "--ngen", "300"], capture_output=True, text=True, timeout=600)
                assert r.returncode == 0
        runs[mode] = [open(re_g.rsplit(".", 1)[0] + ".tour").read() for re_g in refiles]
    assert runs["many"] == runs["single"]


def test_optimize_with_islands_flag(built_lib, tmp_path):
    """`optimize --gpus N`: one population of --npop per GPU, elites exchanged inside the GA kernels.  On a
    1-GPU box the flag is refused by ahx_create (device out of range) like any other bad device."""
    import allhic_b200 as A
    ndev = A._ffi.lib().ahx_device_count()
    ids, clm = _fixture_copy(tmp_path / "i")
    r = subprocess.run([_binary(), "optimize", ids, clm, "--gpus", "2", "--ngen", "3000", "--migrate", "20", "--elite", "1"], capture_output=True, text=True, timeout=300)
    if ndev < 2:
        assert r.returncode == 1 and "device index out of range" in r.stderr
        return
    assert r.returncode == 0 and "Success" in r.stderr and "Island model: 2 populations of 100, 1 elites to every other island every 20 generations" in r.stderr
    last = _parse_last_tour(str(tmp_path / "i" / "test.tour"))
    t = [int(a[3:]) for a, _ in last]
    assert t == list(range(100)) or t == list(range(99, -1, -1))
