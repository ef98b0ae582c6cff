"""GPU: libahx (through the C ABI) against the CPU oracle on identical inputs.
Scores: |gpu - oracle| <= 1e-12 * |oracle| (north_star's fp64 tolerance);
tours / signs / operators: bit-exact."""
import json
import os

import numpy as np
import pytest

from conftest import FIX_CLM, FIX_IDS, GOLDEN
import oracle_lib as O

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def close(got, want, rtol=RTOL):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return bool(np.all(np.abs(got - want) <= rtol * np.abs(want)))


@pytest.fixture(scope="module")
def fix(engine):
    import allhic_b200 as A
    clm = A.CLM(FIX_CLM, FIX_IDS)
    engine.load(clm)
    return engine, clm, O.OracleCLM(FIX_CLM, FIX_IDS)


@pytest.fixture(scope="module")
def syn(built_lib, synth_small):
    import allhic_b200 as A
    ids, clmf, g = synth_small
    eng = A.Engine(0)
    clm = A.CLM(clmf, ids)
    eng.load(clm)
    return eng, clm, O.OracleCLM(clmf, ids)


def rand_signs(rng, n, k=None):
    shape = (n,) if k is None else (k, n)
    return np.where(rng.random(shape) < 0.5, ord('+'), ord('-')).astype(np.uint8)


# ---------------------------------------------------------------- Tour.Evaluate
def test_eval_m_golden_vectors(fix):
    eng, clm, orc = fix
    known = json.load(open(os.path.join(GOLDEN, "known_answers.json")))
    for case in known["cases"]:
        t = np.array(case["tour"], np.int32)
        for kind in ("sparse", "dense"):
            assert close(eng.eval_m(t, kind=kind)[0], case["evaluate"]), (kind, len(t))
        assert close(eng.eval_m(t.astype(np.uint16))[0], case["evaluate"])
        s = np.frombuffer(case["signs"].encode(), np.uint8)
        assert close(eng.eval_q(t, s)[0], case["evaluate_q"])
    assert close(eng.eval_m(np.arange(100, dtype=np.int32))[0], -0.12543807565816803)


@pytest.mark.parametrize("P", [1, 2, 3, 5, 64, 257, 1500])
def test_eval_m_population_fixture(fix, P):
    eng, clm, orc = fix
    rng = np.random.default_rng(P)
    tours = np.stack([rng.permutation(clm.N) for _ in range(P)]).astype(np.int32)
    want = orc.population_evaluate(tours.astype(np.int64), nthreads=4)
    assert close(eng.eval_m(tours), want)
    assert close(eng.eval_m(tours, kind="dense"), want)
    assert close(eng.eval_m(tours.astype(np.uint16)), want)


@pytest.mark.parametrize("path", ["resident", "resident-streamed", "stream32", "fp64"])
@pytest.mark.parametrize("T", ["1", "2", "4"])
def test_eval_m_tours_per_cta(syn, T, path, monkeypatch):
    """All formulations of the contig-space kernel (persistent kernel with the pair
    table resident in shared memory or streamed through a TMA ring; one-CTA-per-batch
    kernel with 32-bit coordinates; fp64 coordinates fallback) at every
    tours-per-batch setting, full tours and sub-tours."""
    eng, clm, orc = syn
    monkeypatch.setenv("AHX_EVAL_T", T)
    monkeypatch.setenv("AHX_EVAL_RES", "1" if path.startswith("resident") else "0")
    monkeypatch.setenv("AHX_EVAL_STREAM", "1" if path == "resident-streamed" else "0")
    monkeypatch.setenv("AHX_EVAL_FP64", "1" if path == "fp64" else "0")
    if path.startswith("resident"):
        assert eng.eval_m_kernel(203) == (3 if path == "resident-streamed" else 2, int(T))
    rng = np.random.default_rng(int(T))
    tours = np.stack([rng.permutation(clm.N) for _ in range(203)]).astype(np.int32)
    tours[0] = np.arange(clm.N); tours[1] = np.arange(clm.N)[::-1]
    want = orc.population_evaluate(tours.astype(np.int64), nthreads=4)
    got = eng.eval_m(tours)
    assert close(got, want)
    assert (got == eng.eval_m(tours)).all()          # fixed-shape reductions: bit-reproducible
    sub = np.stack([rng.permutation(clm.N)[: clm.N // 3] for _ in range(37)]).astype(np.int32)
    assert close(eng.eval_m(sub), orc.population_evaluate(sub.astype(np.int64), nthreads=4))


@pytest.mark.parametrize("n", [2, 3, 9, 31, 32, 33, 65, 129, 250, 513, 1024])
def test_eval_m_random_shapes(built_lib, tmp_path, n):
    """Ragged shapes against the oracle: group sizes around the producers' group widths (8, 32 and 128
    positions), populations that leave batches partly empty, sub-tours of arbitrary length, both index
    widths, tours given in one call and row by row."""
    import allhic_b200 as A
    from allhic_b200 import synth
    g = synth.generate(n, 9000 + n, pairs_per_contig=120 if n > 100 else 400, mean_size=40_000 if n > 300 else 100_000)
    ids, clmf = synth.write(g, str(tmp_path / ("r%d" % n)))
    clm = A.CLM(clmf, ids); eng = A.Engine(0); eng.load(clm); orc = O.OracleCLM(clmf, ids)
    rng = np.random.default_rng(n)
    for P in (1, 3, 4, 7, 150, 593):
        for L in sorted({n, max(1, n - 1), max(1, n // 2), min(n, 5)}):
            tours = np.stack([rng.permutation(n)[:L] for _ in range(P)]).astype(np.int32)
            want = orc.population_evaluate(tours.astype(np.int64), nthreads=4)
            got = eng.eval_m(tours)
            assert close(got, want), (n, P, L)
            assert close(eng.eval_m(tours.astype(np.uint16)), want), (n, P, L, "u16")
            assert ((want == 0) == (got == 0)).all()                # an empty window sum is exactly 0, never -0 or 1e-52
    eng.close()


def test_eval_m_limit_window_and_subtours(syn):
    eng, clm, orc = syn
    n = clm.N
    assert orc.count_W(np.arange(n)) < n * (n - 1) // 2      # the 10 Mb LIMIT truncates
    rng = np.random.default_rng(11)
    for L in (0, 1, 2, 17, n // 2, n - 1):
        tours = np.stack([rng.permutation(n)[:L] for _ in range(9)]).astype(np.int32).reshape(9, L)
        want = orc.population_evaluate(tours.astype(np.int64)) if L else np.zeros(9)
        for kind in ("sparse", "dense"):
            got = eng.eval_m(tours, kind=kind)
            assert close(got, want), (L, kind)
            if L < 2:
                assert (got == 0).all() and not np.signbit(got).any()   # Go starts from +0.0


def test_eval_m_errors(fix, built_lib):
    import allhic_b200 as A
    eng, clm, orc = fix
    bad = np.arange(100, dtype=np.int32); bad[5] = 100
    with pytest.raises(A.AhxError) as ei:
        eng.eval_m(bad)
    assert ei.value.code == -1
    assert close(eng.eval_m(np.arange(100, dtype=np.int32))[0], -0.12543807565816803)   # flag cleared
    e2 = A.Engine(0)
    with pytest.raises(A.AhxError) as ei:
        e2.eval_m(np.arange(4, dtype=np.int32))
    assert ei.value.code == -3
    with pytest.raises(A.AhxError):
        e2.load_arrays([5, 0, 7], [0], [1], [3], [1], -np.ones((1, 8), np.int32), np.zeros((0, 12), np.int32))
    e2.close()


def test_eval_m_giant_group_uses_fp64_coordinates(built_lib, tmp_path):
    """sum(sizes) > 2^31: 2*mid no longer fits 32 bits, the fp64-coordinate kernel must run."""
    import allhic_b200 as A
    rng = np.random.default_rng(13)
    n = 40
    ids = tmp_path / "big.ids"; clmf = tmp_path / "big.clm"
    sizes = rng.integers(1_000_000, 120_000_000, n)
    sizes[:8] = rng.integers(1000, 200_000, 8)
    assert sizes.sum() > 2 ** 31
    ids.write_text("".join("c%d\t1\t%d\n" % (i, s) for i, s in enumerate(sizes)))
    lines = []
    for a in range(n):
        for b in range(a + 1, n):
            if rng.random() < 0.4:
                d = rng.integers(1, 3_000_000, rng.integers(1, 6))
                lines.append("c%d+ c%d-\t%d\t%s\n" % (a, b, len(d), " ".join(map(str, d))))
    clmf.write_text("".join(lines))
    clm = A.CLM(str(clmf), str(ids)); orc = O.OracleCLM(str(clmf), str(ids))
    eng = A.Engine(0); eng.load(clm)
    tours = np.stack([rng.permutation(n) for _ in range(33)]).astype(np.int32)
    tours[:, :8] = np.sort(tours[:, :8], axis=1)
    want = orc.population_evaluate(tours.astype(np.int64))
    assert (want != 0).any()
    assert close(eng.eval_m(tours), want) and close(eng.eval_m(tours, kind="dense"), want)
    eng.close()


# ---------------------------------------------------------------- CLM.EvaluateQ
def test_eval_q_fixture_and_synth(fix, syn):
    for eng, clm, orc in (fix, syn):
        rng = np.random.default_rng(21)
        n = clm.N
        for L in (n, n - 3, n // 3, 2, 1):
            t = rng.permutation(n)[:L].astype(np.int32)
            signs = rand_signs(rng, n, 6)
            signs[0] = ord('+'); signs[1] = ord('-')
            got = eng.eval_q(t, signs)
            orc.set_tour(t)
            for k in range(6):
                orc.set_signs(signs[k])
                assert close(got[k], orc.evaluate_q()), (L, k)
        tours = np.stack([rng.permutation(n) for _ in range(5)]).astype(np.int32)
        signs = rand_signs(rng, n, 5)
        got = eng.eval_q(tours, signs)
        for k in range(5):
            orc.set_tour(tours[k]); orc.set_signs(signs[k])
            assert close(got[k], orc.evaluate_q())


def test_eval_q_global_scratch_path_is_bit_identical(syn, monkeypatch):
    """Groups whose 17 N bytes of per-individual arrays exceed an SM's shared memory (N > ~13 000) keep them in a
    per-CTA slice of a global buffer and walk the individuals grid-stride: same arithmetic, same bits."""
    eng, clm, orc = syn
    rng = np.random.default_rng(33)
    n = clm.N
    tours = np.stack([rng.permutation(n) for _ in range(700)]).astype(np.int32)     # more individuals than CTAs
    signs = rand_signs(rng, n, 700)
    want = eng.eval_q(tours, signs)
    monkeypatch.setenv("AHX_EVALQ_GLOBAL", "1")
    got = eng.eval_q(tours, signs)
    assert (got == want).all()
    sub = np.ascontiguousarray(tours[:5, : n // 2])
    got_sub = eng.eval_q(sub, signs[:5])
    monkeypatch.delenv("AHX_EVALQ_GLOBAL")
    assert (got_sub == eng.eval_q(sub, signs[:5])).all()
    orc.set_tour(tours[3]); orc.set_signs(signs[3])
    assert close(got[3], orc.evaluate_q())


def test_eval_q_general_slot_table(engine, tmp_path):
    """CLM written in both directions with missing orientations: the 8-slot table
    must reproduce the map lookups of Q() (orientation.go:145-150)."""
    import allhic_b200 as A
    rng = np.random.default_rng(4)
    n = 12
    ids = tmp_path / "g.ids"; clmf = tmp_path / "g.clm"
    ids.write_text("".join("c%d\t1\t%d\n" % (i, rng.integers(1000, 900000)) for i in range(n)))
    lines = []
    for _ in range(150):
        a, b = rng.choice(n, 2, replace=False)
        ao, bo = rng.choice(list("+-"), 2)
        d = rng.integers(1, 3_000_000, rng.integers(1, 9))
        lines.append("c%d%s c%d%s\t%d\t%s\n" % (a, ao, b, bo, len(d), " ".join(map(str, d))))
    clmf.write_text("".join(lines))
    clm = A.CLM(str(clmf), str(ids)); orc = O.OracleCLM(str(clmf), str(ids))
    eng = A.Engine(0); eng.load(clm)
    for _ in range(20):
        t = rng.permutation(n)[: rng.integers(2, n + 1)].astype(np.int32)
        s = rand_signs(rng, n)
        orc.set_tour(t); orc.set_signs(s)
        assert close(eng.eval_q(t, s)[0], orc.evaluate_q(literal=True))
        assert close(eng.eval_m(t)[0], orc.evaluate(t.astype(np.int64)))
        assert close(eng.eval_m(t, kind="dense")[0], orc.evaluate(t.astype(np.int64)))
    eng.close()


def test_group_without_contacts_and_heavy_pairs(built_lib, tmp_path):
    """Edge inputs the reference accepts: (1) a CLM that names none of the REfile's contigs --
    M is all zero, every score is +0.0 (Go's score starts from +0.0 and nothing is subtracted),
    the GA still returns a permutation, flips REJECT; (2) pairs with hundreds / thousands of links
    (the resident table stores 8-bit counts and splits such pairs over several entries)."""
    import allhic_b200 as A
    rng = np.random.default_rng(12)
    n = 40
    ids = tmp_path / "e.ids"; clmf = tmp_path / "e.clm"
    ids.write_text("".join("c%d\t1\t%d\n" % (i, rng.integers(1000, 90000)) for i in range(n)))
    clmf.write_text("x1+ x2+\t3\t5 6 7\n")                                # unknown contigs: skipped (clm.go:211-218)
    clm = A.CLM(str(clmf), str(ids)); eng = A.Engine(0); eng.load(clm)
    tours = np.stack([rng.permutation(n) for _ in range(7)]).astype(np.int32)
    for kind in ("sparse", "dense"):
        s = eng.eval_m(tours, kind=kind)
        assert (s == 0).all() and not np.signbit(s).any()
    signs = rand_signs(rng, n)
    assert eng.eval_q(tours[0], signs)[0] == 0
    best, st = eng.ga_run(tours[0], eng.ga_params(npop=32, ngen=20, max_gens=200, seed=1))
    assert sorted(best.tolist()) == list(range(n)) and st.best_score == 0 and st.generations == 21
    s2, acc, a, b = eng.flip_whole(tours[0], signs)
    assert not acc and (s2 == signs).all()                                 # newScore <= score: REJECT (orientation.go:78)
    s3, any_acc, na, nr, *_ = eng.flip_one(tours[0], signs)
    assert not any_acc and na == 0 and nr == n and (s3 == signs).all()
    eng.close()
    # (2) heavy pairs: 255, 256, 700 and 2600 links
    lines = []
    for (a, b), cnt in zip([(0, 1), (1, 2), (2, 5), (3, 9)], [255, 256, 700, 2600]):
        d = rng.integers(1, 200000, cnt)
        for ao, bo in ("++", "+-", "-+", "--"):
            lines.append("c%d%s c%d%s\t%d\t%s\n" % (a, ao, b, bo, cnt, " ".join(map(str, d))))
    for _ in range(60):
        a, b = sorted(rng.choice(n, 2, replace=False))
        d = rng.integers(1, 200000, rng.integers(3, 12))
        lines.append("c%d+ c%d+\t%d\t%s\n" % (a, b, len(d), " ".join(map(str, d))))
    clmf.write_text("".join(lines))
    clm = A.CLM(str(clmf), str(ids)); orc = O.OracleCLM(str(clmf), str(ids))
    eng = A.Engine(0); eng.load(clm)
    assert clm.tables()["nlinks"].max() == 2600
    want = orc.population_evaluate(tours.astype(np.int64))
    assert eng.eval_m_kernel(len(tours))[0] == 2                           # the resident-table kernel is the one under test
    assert close(eng.eval_m(tours), want) and close(eng.eval_m(tours, kind="dense"), want)
    best, st = eng.ga_run(tours[0], eng.ga_params(npop=64, ngen=30, max_gens=300, seed=2))
    assert close(st.best_score, -orc.evaluate(best.astype(np.int64)))
    eng.close()


# ---------------------------------------------------------------- operators
def test_mutation_operators_bit_exact(fix):
    eng, clm, orc = fix
    rng = np.random.default_rng(31)
    for L in (2, 3, 10, 100):
        base = rng.permutation(100)[:L].astype(np.int32)
        for _ in range(25):
            p, q = sorted(rng.integers(0, L, 2).tolist())
            assert (eng.mutate(base, 0, p, q) == O.mutate(base, "permute", p=p, q=q)).all()
            assert (eng.mutate(base, 3, p, q) == O.mutate(base, "inversion", p=p, q=q)).all()
            for d in (0, 1):
                assert (eng.mutate(base, 2, p, q, d) == O.mutate(base, "insertion", p=p, q=q, front=d)).all()
            k = int(rng.integers(1, L))
            assert (eng.mutate(base, 1, k) == O.mutate(base, "splice", k=k)).all()


# ---------------------------------------------------------------- GARun
def test_ga_run_validity_and_bookkeeping(fix):
    eng, clm, orc = fix
    rng = np.random.default_rng(41)
    init = rng.permutation(clm.N).astype(np.int32)
    f0 = -orc.evaluate(init.astype(np.int64))
    seen = []
    prm = eng.ga_params(npop=100, ngen=40, max_gens=100000, seed=42, log_every=50, phase=2)
    best, st = eng.ga_run(init, prm, callback=lambda ph, g, s, t: seen.append((ph, g, s, sorted(t.tolist()) == list(range(clm.N)))))
    assert sorted(best.tolist()) == list(range(clm.N))                       # bit-exact validity
    assert close(st.best_score, -orc.evaluate(best.astype(np.int64)))        # HoF score is the tour's score
    assert st.best_score >= f0 and st.stopped == 1
    assert st.generations - st.updated == 41                                 # EarlyStop: gen - updated > NGen
    assert seen[0][:2] == (2, 0) and close(seen[0][2], f0)
    assert [g for _, g, _, _ in seen] == list(range(0, st.generations + 1, 50))
    assert all(ok for *_, ok in seen) and all(b[2] >= a[2] for a, b in zip(seen, seen[1:]))
    assert st.evaluations > 0.1 * 100 * st.generations * 0.5
    best2, st2 = eng.ga_run(init, prm)                                       # same seed => same run
    assert (best2 == best).all() and st2.best_score == st.best_score and st2.generations == st.generations
    best3, st3 = eng.ga_run(init, eng.ga_params(npop=100, ngen=40, max_gens=100000, seed=43))
    assert st3.generations != st.generations or (best3 != best).any()
    _, st4 = eng.ga_run(init, eng.ga_params(npop=64, ngen=5000, max_gens=37, seed=1))
    assert st4.generations == 37 and st4.stopped == 1


def test_ga_matches_or_beats_cpu_restatement(fix):
    """Final GA tours must score at least as well as the CPU restatement's GA
    over seeds 42..46 (same npop / patience / mutation rate)."""
    eng, clm, orc = fix
    ours, theirs = [], []
    for seed in range(42, 47):
        orc.activate(shuffle=True, seed=seed, flip_all=False)
        init = orc.tour().astype(np.int32)
        r = orc.ga_run(npop=100, ngen=300, mutprob=0.2, seed=seed, max_gens=100000, nthreads=4)
        _, st = eng.ga_run(init, eng.ga_params(npop=100, ngen=300, seed=seed))
        ours.append(st.best_score); theirs.append(r.best_score)
    assert np.median(ours) >= 0.98 * np.median(theirs), (ours, theirs)
    # with the population the GPU is built for, never worse than the CPU run
    _, st = eng.ga_run(init, eng.ga_params(npop=4096, ngen=300, seed=46))
    assert st.best_score >= max(theirs) * 0.999, (st.best_score, theirs)


def test_ga_elites_and_migrants(fix):
    import allhic_b200 as A
    eng, clm, orc = fix
    rng = np.random.default_rng(51)
    init = rng.permutation(clm.N).astype(np.int32)
    eng.ga_begin(init, eng.ga_params(npop=128, ngen=100000, seed=5))
    st = eng.ga_advance(60)
    assert st.generations == 60 and not st.stopped
    tours, fit = eng.ga_get_elites(4)
    assert (np.diff(fit) >= 0).all() and close(-fit[0], st.best_score) or -fit[0] <= st.best_score
    for t, f in zip(tours, fit):
        assert sorted(t.tolist()) == list(range(clm.N)) and close(f, orc.evaluate(t.astype(np.int64)))
    ident = np.arange(clm.N, dtype=np.int32)[None, :]
    eng.ga_put_migrants(ident)                                  # the true order beats anything found in 60 gens
    bt, bs = eng.ga_best()
    assert (bt == ident[0]).all() and close(bs, 0.12543807565816803)
    with pytest.raises(A.AhxError):
        eng.ga_put_migrants(np.zeros((1, clm.N), np.int32))     # not a permutation
    st = eng.ga_advance(10)
    assert st.generations == 70 and st.best_score >= bs
    eng.ga_end()


@pytest.mark.parametrize("path", ["persistent", "persistent-streamed", "two-kernel", "fp64", "global"])
@pytest.mark.parametrize("gaT", ["1", "2", "4"])
def test_ga_population_is_consistent_on_every_path(syn, path, gaT, monkeypatch):
    """After many generations every individual must still be what its stored state says:
    a permutation of the active contigs, fitness == a fresh Tour.Evaluate of its tour
    (1e-12), and -- persistent kernel -- stored coordinates == the tour's coordinates
    (they are derived by per-operator transforms, never recomputed).  Same seed => the
    three implementations of CLM.GARun follow the same trajectory."""
    eng, clm, orc = syn
    if not path.startswith("persistent"):
        monkeypatch.setenv("AHX_GA_RES", "0")
    monkeypatch.setenv("AHX_EVAL_STREAM", "1" if path == "persistent-streamed" else "0")
    if path == "fp64":
        monkeypatch.setenv("AHX_EVAL_FP64", "1")
        monkeypatch.setenv("AHX_EVAL_RES", "0")
    if path == "global":        # what groups beyond ~26 000 contigs run: rows and coordinates in global memory
        if gaT != "1":
            pytest.skip("one variant")
        monkeypatch.setenv("AHX_GA_GLOBAL", "1")
    monkeypatch.setenv("AHX_GA_T", gaT)
    rng = np.random.default_rng(77)
    init = rng.permutation(clm.N).astype(np.int32)
    eng.ga_begin(init, eng.ga_params(npop=600, ngen=1 << 30, max_gens=1 << 40, seed=9, mut_prob=0.35))
    st = eng.ga_advance(120)
    assert st.generations == 120 and eng.ga_selfcheck() == 0
    tours, fit = eng.ga_get_elites(3)
    eng.ga_put_migrants(np.stack([rng.permutation(clm.N) for _ in range(5)]).astype(np.int32))
    assert eng.ga_selfcheck() == 0
    st = eng.ga_advance(37)
    assert st.generations == 157 and eng.ga_selfcheck() == 0
    best, score = eng.ga_best()
    eng.ga_end()
    assert close(score, -orc.evaluate(best.astype(np.int64)))
    key = "ga_traj"
    ref = test_ga_population_is_consistent_on_every_path.__dict__.setdefault(key, {})
    ref.setdefault("score", score)
    assert close(score, ref["score"], 1e-9)          # same draws, same decisions (scores agree to rounding)


def test_ga_mutation_heavy_and_tiny_populations(fix):
    """mut_prob = 1 needs a fresh pool row for every offspring; npop = 4 is the smallest
    population two tournaments of 3 allow."""
    eng, clm, orc = fix
    rng = np.random.default_rng(5)
    init = rng.permutation(clm.N).astype(np.int32)
    for npop, mp in ((64, 1.0), (4, 0.5), (1000, 0.0)):
        eng.ga_begin(init, eng.ga_params(npop=npop, ngen=1 << 30, max_gens=1 << 40, seed=3, mut_prob=mp))
        st = eng.ga_advance(200)
        assert st.generations == 200 and eng.ga_selfcheck() == 0
        best, score = eng.ga_best()
        assert close(score, -orc.evaluate(best.astype(np.int64)))
        if mp == 0.0:
            assert (best == init).all() and st.evaluations == 1
        eng.ga_end()


@pytest.mark.parametrize("n", [5, 33, 65, 129, 250])
def test_ga_ragged_group_sizes(built_lib, tmp_path, n):
    """The persistent GA kernel on group sizes that leave its 32/T-contig groups ragged, with small odd
    populations: every individual stays consistent with its stored tour, the best matches the oracle."""
    import allhic_b200 as A
    from allhic_b200 import synth
    g = synth.generate(n, 9100 + n, pairs_per_contig=400)
    ids, clmf = synth.write(g, str(tmp_path / ("g%d" % n)))
    clm = A.CLM(clmf, ids); eng = A.Engine(0); eng.load(clm); orc = O.OracleCLM(clmf, ids)
    rng = np.random.default_rng(n)
    for npop, mp in ((10, 0.5), (37, 0.2), (301, 0.9)):
        init = rng.permutation(n).astype(np.int32)
        eng.ga_begin(init, eng.ga_params(npop=npop, ngen=1 << 30, max_gens=1 << 40, seed=n + npop, mut_prob=mp))
        st = eng.ga_advance(60)
        assert st.generations == 60 and eng.ga_selfcheck() == 0, (n, npop)
        best, score = eng.ga_best()
        assert sorted(best.tolist()) == list(range(n)) and close(score, -orc.evaluate(best.astype(np.int64)))
        eng.ga_end()
    eng.close()


def test_ga_large_population_stays_on_the_persistent_kernel(fix):
    """Populations beyond one mask word per thread (2P > 32 * 768) run the chunked block scans."""
    eng, clm, orc = fix
    rng = np.random.default_rng(8)
    init = rng.permutation(clm.N).astype(np.int32)
    for npop in (12289, 40000):
        eng.ga_begin(init, eng.ga_params(npop=npop, ngen=1 << 30, max_gens=1 << 40, seed=9))
        assert eng.ga_kernel() in (2, 3)
        st = eng.ga_advance(40)
        assert st.generations == 40 and eng.ga_selfcheck() == 0
        best, score = eng.ga_best()
        assert close(score, -orc.evaluate(best.astype(np.int64)))
        eng.ga_end()


def test_optimize_many_equals_separate_runs(built_lib, tmp_path):
    """`allhic-b200 optimize-many` (one process and one device context for a list of groups, the host
    binary's form of pipeline's loop, allhic.go:285-293) leaves the same .tour files as separate
    `optimize` runs; the second group goes through a context that has already held another one."""
    import shutil, subprocess
    from allhic_b200 import farm, synth
    g = synth.generate(120, 4242)
    runs = {}
    for mode in ("many", "single"):
        d = tmp_path / mode
        (d / "a").mkdir(parents=True); (d / "b").mkdir()
        ids_a, clm_a = synth.write(g, str(d / "a" / "s120"))
        shutil.copy(FIX_IDS, d / "b" / "test.ids"); shutil.copy(FIX_CLM, d / "b" / "test.clm")
        pairs = [(ids_a, clm_a), (str(d / "b" / "test.ids"), str(d / "b" / "test.clm"))]
        if mode == "many":
            (d / "list").write_text("".join("%s %s\n" % p for p in pairs))
            r = subprocess.run([farm.BINARY, "optimize-many", str(d / "list"), "--ngen", "300"], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0 and r.stderr.count("Success") == 2
            assert [l.split()[1] for l in r.stdout.splitlines() if l.startswith("AHX_GROUP_DONE")] == ["0", "1"]
        else:
            for p in pairs:
                r = subprocess.run([farm.BINARY, "optimize", p[0], p[1], "--ngen", "300"], capture_output=True, text=True, timeout=300)
                assert r.returncode == 0 and "Success" in r.stderr
        runs[mode] = [open(p[0].rsplit(".", 1)[0] + ".tour").read() for p in pairs]
    assert runs["many"] == runs["single"]


# ---------------------------------------------------------------- orientation
def test_flip_whole_and_one_match_oracle(fix, syn):
    for eng, clm, orc in (fix, syn):
        rng = np.random.default_rng(61)
        n = clm.N
        for L in (n, n // 2):
            t = rng.permutation(n)[:L].astype(np.int32)
            s = rand_signs(rng, n)
            orc.set_tour(t); orc.set_signs(s)
            acc_o, a_o, b_o = orc.flip_whole()
            s1, acc, a, b = eng.flip_whole(t, s)
            assert acc == acc_o and close(a, a_o) and close(b, b_o) and (s1 == orc.signs()).all()
            any_o, na_o, nr_o, fs_o, trace_o, steps_o = orc.flip_one()
            s2, any_, na, nr, fs, trace, steps = eng.flip_one(t, s1)
            assert (steps == steps_o).all() and (na, nr, any_) == (na_o, nr_o, any_o)
            assert (s2 == orc.signs()).all()                                  # bit-exact orientations
            assert close(fs, fs_o, 1e-11) and close(trace, trace_o, 1e-11)
            assert close(eng.eval_q(t, s2)[0], fs_o)


def test_flip_one_rejects_foreign_sign_bytes(fix):
    """CLM.Signs only ever holds '+' / '-' (Activate writes '+', every change goes through rr,
    clm.go:160-165); ahx_flip_one refuses anything else instead of imitating the reference's
    stale-score behaviour on such input.  EvaluateQ itself accepts it (no stored orientation
    matches, the pair is skipped, orientation.go:143)."""
    import allhic_b200 as A
    eng, clm, orc = fix
    rng = np.random.default_rng(62)
    t = rng.permutation(clm.N).astype(np.int32)
    weird = rand_signs(rng, clm.N); weird[t[3]] = ord('?')
    with pytest.raises(A.AhxError) as ei:
        eng.flip_one(t, weird)
    assert ei.value.code == -1
    orc.set_tour(t); orc.set_signs(weird)
    assert close(eng.eval_q(t, weird)[0], orc.evaluate_q())


def test_flip_all_matches_dense_eigensolver(fix):
    eng, clm, orc = fix
    rng = np.random.default_rng(71)
    t = rng.permutation(clm.N).astype(np.int32)
    plus = np.full(clm.N, ord('+'), np.uint8)
    orc.set_tour(t); orc.set_signs(plus)
    acc_o, a_o, b_o = orc.flip_all()
    s, acc, a, b = eng.flip_all(t, plus)
    assert close(a, a_o)
    so = orc.signs()
    # the eigenvector's global sign is arbitrary: same pattern or its complement
    same = (s == so).all()
    comp = acc and acc_o and (s != so).all()
    if not (same or comp):
        # near-degenerate top eigenvalues: accept any pattern that scores as well
        assert b >= b_o - 1e-9 * abs(b_o)


@pytest.mark.parametrize("n", [30, 400])
def test_flip_all_sign_pattern_vs_jacobi_oracle(built_lib, tmp_path, n):
    """flipAll at sizes on both sides of the dense/Lanczos switch (<= 48 contigs: dense Jacobi like the reference's
    mat64.EigenSym; above: sparse Lanczos): the sign pattern against the oracle's dense Jacobi decomposition of O
    (orientation.go:31-64), strand-mixed contacts, and the solver's own convergence report."""
    import allhic_b200 as A
    from allhic_b200 import synth
    g = synth.generate(n, 9500 + n, pairs_per_contig=200)
    ids, clmf = synth.write(g, str(tmp_path / ("o%d" % n)))
    # flip a third of the contigs: rewrite the CLM with swapped orientation tags for those contigs, so that the
    # kept contact of a pair (smallest SumLog, clm.go:229-236) has mixed strandedness and O has negative entries
    rng = np.random.default_rng(n)
    flipped = set(np.flatnonzero(rng.random(n) < 0.33).tolist())
    rr = {"+": "-", "-": "+"}
    out = []
    for row in open(clmf):
        w = row.rstrip("\n").split("\t")
        at, bt = w[0].split(" ")
        a, b = int(at[3:-1]), int(bt[3:-1])
        ao = rr[at[-1]] if a in flipped else at[-1]
        bo = rr[bt[-1]] if b in flipped else bt[-1]
        out.append("%s%s %s%s\t%s\t%s\n" % (at[:-1], ao, bt[:-1], bo, w[1], w[2]))
    open(clmf, "w").write("".join(out))
    clm = A.CLM(clmf, ids); eng = A.Engine(0); eng.load(clm); orc = O.OracleCLM(clmf, ids)
    assert (clm.tables()["strand"] < 0).any() and (clm.tables()["strand"] > 0).any()
    t = rng.permutation(n).astype(np.int32)
    plus = np.full(n, ord('+'), np.uint8)
    orc.set_tour(t); orc.set_signs(plus)
    acc_o, a_o, b_o = orc.flip_all()
    s, acc, a, b = eng.flip_all(t, plus)
    info = eng.flip_all_info()
    assert info["converged"] and info["dense"] == (n <= 48) and info["residual"] <= 1e-9
    assert close(a, a_o)
    # The eigenvector's global sign is implementation defined (gonum's EigenSym, the oracle's Jacobi and the
    # product's Lanczos may each return v or -v), so the candidate sign vector is the oracle's pattern or its
    # complement; the new score, and with it ACCEPT / REJECT (orientation.go:56-62), belongs to the candidate.
    v = orc.top_eigvec()[0]
    cand = np.where(v < 0, ord('-'), ord('+')).astype(np.uint8)
    comp = np.where(v < 0, ord('+'), ord('-')).astype(np.uint8)
    orc.set_signs(cand); q_cand = orc.evaluate_q()
    orc.set_signs(comp); q_comp = orc.evaluate_q()
    assert close(b_o, q_cand)
    assert close(b, q_cand) or close(b, q_comp)
    ours = cand if close(b, q_cand) else comp
    assert acc == (not b < a)                               # REJECT only when newScore < score (orientation.go:58)
    # contigs without any contact have a zero component: '+' on both sides whatever the global sign (`v < 0`)
    live = np.abs(v) > 1e-9 * np.abs(v).max()
    assert live.sum() >= 0.9 * n
    assert (s[live] == (ours if acc else plus)[live]).all()
    if close(b, q_cand):
        assert acc == acc_o
    eng.close()


def test_flip_all_self_pair_lines_reach_the_diagonal_of_O(built_lib, tmp_path):
    """A CLM line naming one contig twice lands on the diagonal of O in the reference (SetSym(a, a, ...),
    orientation.go:124-132) and so changes flipAll's eigenvector; M() and Q() never read it."""
    import allhic_b200 as A
    rng = np.random.default_rng(3)
    n = 9
    ids = tmp_path / "d.ids"; clmf = tmp_path / "d.clm"
    ids.write_text("".join("c%d\t1\t%d\n" % (i, rng.integers(1000, 90000)) for i in range(n)))
    lines = []
    for a in range(n):
        for b in range(a + 1, n):
            if rng.random() < 0.6:
                d = rng.integers(1, 200000, rng.integers(3, 9))
                lines.append("c%d%s c%d%s\t%d\t%s\n" % (a, "+-"[rng.integers(2)], b, "+-"[rng.integers(2)], len(d), " ".join(map(str, d))))
    base = "".join(lines)
    t = rng.permutation(n).astype(np.int32)
    plus = np.full(n, ord('+'), np.uint8)
    pats = []
    for extra in ("", "c2+ c2-\t400\t%s\n" % " ".join(["5000"] * 400)):
        clmf.write_text(base + extra)
        clm = A.CLM(str(clmf), str(ids)); eng = A.Engine(0); eng.load(clm); orc = O.OracleCLM(str(clmf), str(ids))
        orc.set_tour(t); orc.set_signs(plus)
        orc.flip_all()
        s, acc, a, b = eng.flip_all(t, plus)
        so = orc.signs()
        assert (s == so).all() or (s != so).all(), extra[:8]
        assert close(eng.eval_m(t)[0], orc.evaluate(t.astype(np.int64)))          # M() never reads the diagonal
        pats.append(orc.top_eigvec()[0].copy())
        eng.close()
    assert not np.allclose(np.abs(pats[0]), np.abs(pats[1]))   # the diagonal entry did change the eigenvector


# ---------------------------------------------------------------- full-size properties
def test_config3_size_properties(built_lib):
    """N=2000 / P=8000 (BASELINE config 3): size-independent properties instead of
    a full oracle run: sparse == dense kernel, reversal symmetry, determinism, and
    an oracle spot check on a sample."""
    import allhic_b200 as A
    from allhic_b200 import synth
    ids, clmf, _ = synth.materialize("config3")
    clm = A.CLM(clmf, ids); eng = A.Engine(0); eng.load(clm)
    rng = np.random.default_rng(81)
    P, n = 8000, clm.N
    tours = np.stack([rng.permutation(n) for _ in range(P)]).astype(np.int32)
    s = eng.eval_m(tours)
    assert close(eng.eval_m(tours[:512], kind="dense"), s[:512])
    assert close(eng.eval_m(tours[:, ::-1].copy()), s, 1e-13)        # |x_a - x_b| is reversal invariant
    assert close(eng.eval_m(tours.astype(np.uint16)), s, 1e-13)     # other batch shape: last bits may differ
    orc = O.OracleCLM(clmf, ids)
    idx = rng.choice(P, 24, replace=False)
    assert close(s[idx], orc.population_evaluate(tours[idx].astype(np.int64), nthreads=8))
    sg = rand_signs(rng, n, 3)
    q = eng.eval_q(tours[0], sg)
    orc.set_tour(tours[0])
    for k in range(3):
        orc.set_signs(sg[k]); assert close(q[k], orc.evaluate_q())
    assert (s == eng.eval_m(tours)).all()                           # bit-reproducible run to run
    # the pair table is the pair list rounded up to one loop iteration of the consumers (2048 entries):
    # lanes without a conflict-free candidate take colliding pairs, not padding (order_pairs)
    nl = clm.tables()["nlinks"]
    need = int(np.sum((nl + 254) // 255))
    assert [eng.pair_table_entries(t) for t in (1, 2, 4)] == [(need + 2047) // 2048 * 2048] * 3
    # invalid entries in the producers' short path (whole 16-position groups inside the range): any of
    # -1, N, a huge index, in the middle of a tour of the middle of the population, is an argument error
    import allhic_b200 as A
    for badv in (-1, n, 1 << 30):
        t2 = tours[:64].copy(); t2[37, 1001] = badv
        with pytest.raises(A.AhxError) as ei:
            eng.eval_m(t2)
        assert ei.value.code == -1
    assert close(eng.eval_m(tours[:64]), s[:64], 1e-13)              # flag cleared (64 tours: another T, another summation order)
    # the device-resident GA at the full size: every one of the 8000 individuals stays consistent with
    # its stored tour (permutation, fitness, coordinates) and the best-ever score never decreases
    eng.ga_begin(tours[0], eng.ga_params(npop=P, ngen=1 << 30, max_gens=1 << 40, seed=11))
    last = -np.inf
    for _ in range(3):
        st = eng.ga_advance(40)
        assert st.best_score >= last
        last = st.best_score
    assert st.generations == 120 and eng.ga_selfcheck() == 0
    best, score = eng.ga_best()
    assert sorted(best.tolist()) == list(range(n)) and close(score, -orc.evaluate(best.astype(np.int64)))
    assert score > -s[0] * 1.5                                      # 120 generations improve a random tour a lot
    eng.ga_end()
    eng.close()
