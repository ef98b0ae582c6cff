"""GPU: randomized stress of the persistent GA kernel's generation loop (ahx_ga_res.cuh).  The loop starts the next
generation's selection chains behind the grid barrier while TMA bulk copies land its mask / row bit map / 128-bit
minimum, hands out pool rows from a bit map and settles the hall of fame with a 128-bit atomic minimum -- all of it
invisible in the results as long as it is right.  So: many (population, mutation rate, launch chunking, tournament
size, island) combinations, each checked three ways:
  * every individual is what its stored state says (ahx_ga_selfcheck: permutation, fitness == fresh Tour.Evaluate
    to 1e-12, stored coordinates == the tour's);
  * the run is bit-reproducible and does not depend on how the generations are cut into launches;
  * the hall of fame is the best individual ever evaluated: its score equals the oracle's score of its tour and the
    best-ever trace never decreases.
"""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def close(got, want, rtol=1e-12):
    return abs(got - want) <= rtol * abs(want)


@pytest.fixture(scope="module")
def syn(built_lib, synth_small):
    import allhic_b200 as A
    ids, clmf, g = synth_small
    eng = A.Engine(0)
    clm = A.CLM(clmf, ids)
    eng.load(clm)
    return eng, clm, O.OracleCLM(clmf, ids)


def _run(eng, init, prm_kw, chunks):
    eng.ga_begin(init, eng.ga_params(**prm_kw))
    trace = []
    st = None
    for c in chunks:
        st = eng.ga_advance(c)
        trace.append(st.best_score)
    bad = eng.ga_selfcheck()
    best, score = eng.ga_best()
    kernel = eng.ga_kernel()
    eng.ga_end()
    return best, score, st, trace, bad, kernel


@pytest.mark.parametrize("case", range(int(os.environ.get("AHX_STRESS_CASES", "24"))))   # AHX_STRESS_CASES=200 for a long soak
def test_generation_loop_random_configurations(syn, case):
    eng, clm, orc = syn
    rng = np.random.default_rng(9000 + case)
    n = clm.N
    npop = int(rng.choice([4, 5, 33, 100, 257, 1000, 2049, 8191, 8193, 12000]))
    mut_prob = float(rng.choice([0.0, 0.03, 0.2, 0.5, 0.97, 1.0]))
    k = int(rng.choice([3, 3, 3, 2, 5])) if npop >= 8 else 3
    total = int(rng.integers(20, 90)) if npop <= 3000 else int(rng.integers(8, 25))
    cuts = np.sort(rng.choice(np.arange(1, total), size=min(total - 1, int(rng.integers(0, 6))), replace=False)).tolist()
    chunks = np.diff([0] + cuts + [total]).tolist()
    init = rng.permutation(n).astype(np.int32)
    prm = dict(npop=npop, ngen=1 << 30, max_gens=1 << 40, seed=int(rng.integers(1, 1 << 30)), mut_prob=mut_prob, tournament_k=k)
    best, score, st, trace, bad, kernel = _run(eng, init, prm, chunks)
    assert kernel >= 2                                           # the persistent kernel ran
    assert st.generations == total and bad == 0, (npop, mut_prob, k, chunks)
    assert sorted(best.tolist()) == list(range(n))
    assert close(score, -orc.evaluate(best.astype(np.int64)))
    assert all(b >= a for a, b in zip(trace, trace[1:]))
    if mut_prob == 0.0:
        assert (best == init).all() and st.evaluations == 1
    # one launch instead of several, and a second identical run: the same bits
    best1, score1, st1, _, bad1, _ = _run(eng, init, prm, [total])
    assert bad1 == 0 and (best1 == best).all() and score1 == score and st1.evaluations == st.evaluations


def test_early_stop_generation_is_independent_of_launch_chunking(syn):
    """EarlyStop (evaluate.go:304-306) fires at the same generation whether the run is one launch or many -- the
    loop's speculative front half must leave no trace when the run stops."""
    eng, clm, orc = syn
    init = np.random.default_rng(4).permutation(clm.N).astype(np.int32)
    prm = dict(npop=64, ngen=40, max_gens=100000, seed=11)
    eng.ga_begin(init, eng.ga_params(**prm))
    a = eng.ga_advance(100000)
    ta, sa = eng.ga_best()
    assert a.stopped == 1 and eng.ga_selfcheck() == 0
    eng.ga_end()
    eng.ga_begin(init, eng.ga_params(**prm))
    b = None
    for _ in range(100000):
        b = eng.ga_advance(3)
        if b.stopped:
            break
    tb, sb = eng.ga_best()
    assert eng.ga_selfcheck() == 0
    eng.ga_end()
    assert b.generations == a.generations and b.updated == a.updated and sb == sa and (ta == tb).all()
    # ... and a stopped run can be inspected and restarted by migrants without corrupting the population
    eng.ga_begin(init, eng.ga_params(**prm))
    eng.ga_advance(100000)
    eng.ga_put_migrants(np.arange(clm.N, dtype=np.int32)[None, :])
    assert eng.ga_selfcheck() == 0
    c = eng.ga_advance(5)
    assert eng.ga_selfcheck() == 0 and c.generations >= a.generations
    eng.ga_end()


@pytest.mark.parametrize("W,npop,ne,K", [(2, 40, 1, 3), (3, 64, 2, 5), (4, 33, 3, 4)])
def test_islands_random_chunking_is_reproducible(built_lib, synth_small, W, npop, ne, K):
    import threading
    import allhic_b200 as A
    ids, clmf, g = synth_small
    clm = A.CLM(clmf, ids)
    ndev = A._ffi.lib().ahx_device_count()
    engs = [A.Engine(i % ndev) for i in range(W)]
    for e in engs:
        e.load(clm)
    init = np.random.default_rng(W).permutation(clm.N).astype(np.int32)

    def run(chunks):
        for i, e in enumerate(engs):
            e.ga_begin(init, e.ga_params(npop=npop, ngen=1 << 30, max_gens=1 << 40, seed=77, island=i, n_islands=W, migrate_every=K, n_elite=ne))
        A.Engine.ga_island_connect_local(engs)
        for c in chunks:
            th = [threading.Thread(target=e.ga_advance, args=(c,)) for e in engs]
            for t in th:
                t.start()
            for t in th:
                t.join()
        out = []
        for e in engs:
            assert e.ga_selfcheck() == 0
            out.append((e.ga_best()[0].copy(), e.ga_best()[1], e.ga_island_epochs()))
            e.ga_end()
        return out
    a = run([37])
    b = run([K, 1, K - 1, 2 * K + 1, 37 - 4 * K - 1])
    for (ta, sa, ea), (tb, sb, eb) in zip(a, b):
        assert (ta == tb).all() and sa == sb and ea == eb == 37 // K
    for e in engs:
        e.close()
