"""GPU: the island model behind the C ABI (include/ahx.h "Islands").  The exchange step runs
inside the persistent GA kernel over peer memory; here it is checked against what it must do:
migrants are exactly the other islands' elites (as a twin single-population run of the same
Philox stream produces them), they replace the worst individuals, every island stays
consistent with its stored tours, the collective EarlyStop fires on all islands at the same
generation, and the run is bit-reproducible.

Islands of one process may share a GPU as long as their cooperative grids fit side by side
(small populations), which is how this runs on a 1-GPU box; with >= 2 GPUs the same tests
spread the islands over them, and the CUDA-IPC wiring (one process per GPU) is tested too."""
import os
import socket
import threading

import numpy as np
import pytest

from conftest import FIX_CLM, FIX_IDS
import oracle_lib as O

pytestmark = pytest.mark.gpu


def close(got, want, rtol=1e-12):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return bool(np.all(np.abs(got - want) <= rtol * np.abs(want)))


def _engines(n, clm):
    import allhic_b200 as A
    ndev = A._ffi.lib().ahx_device_count()
    engs = [A.Engine(i % ndev) for i in range(n)]
    for e in engs:
        e.load(clm)
    return engs


def _advance_all(engs, n_gens):
    """ga_advance on every island at once (their kernels wait for each other): one thread per island."""
    out, errs = [None] * len(engs), []

    def work(i):
        try:
            out[i] = engs[i].ga_advance(n_gens)
        except Exception as e:  # pragma: no cover
            errs.append(e)
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(engs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise errs[0]
    return out


@pytest.fixture(scope="module")
def clm(built_lib):
    import allhic_b200 as A
    return A.CLM(FIX_CLM, FIX_IDS)


@pytest.mark.parametrize("W,ne", [(2, 2), (3, 1), (4, 3)])
def test_first_exchange_moves_exactly_the_elites(clm, W, ne):
    import allhic_b200 as A
    orc = O.OracleCLM(FIX_CLM, FIX_IDS)
    K, P = 7, 48
    init = np.random.default_rng(3).permutation(clm.N).astype(np.int32)
    engs = _engines(W, clm)
    prm = lambda i, w: engs[0].ga_params(npop=P, ngen=1 << 30, max_gens=1 << 40, seed=21, island=i, n_islands=w, migrate_every=K, n_elite=ne)
    # twins: the same Philox streams as single populations, K generations = the state just before the exchange
    twin = A.Engine(0); twin.load(clm)
    pre = []
    for i in range(W):
        twin.ga_begin(init, prm(i, 1))
        twin.ga_advance(K)
        tours, fit = twin.ga_get_elites(P)
        pre.append((tours.copy(), fit.copy()))
        twin.ga_end()
    twin.close()
    for i, e in enumerate(engs):
        e.ga_begin(init, prm(i, W))
        assert e.ga_kernel() == 2
    with pytest.raises(A.AhxError):
        engs[0].ga_advance(1)                                  # not connected yet
    A.Engine.ga_island_connect_local(engs)
    sts = _advance_all(engs, K)
    m = ne * (W - 1)
    for i, e in enumerate(engs):
        assert sts[i].generations == K and e.ga_island_epochs() == 1 and e.ga_selfcheck() == 0
        tours, fit = e.ga_get_elites(P)
        # expected population: own pre-exchange population minus its m worst, plus the others' ne best
        own_t, own_f = pre[i]
        want_f = sorted(list(own_f[: P - m]) + [f for j in range(W) if j != i for f in pre[j][1][:ne]])
        assert np.array_equal(np.sort(fit), np.array(want_f)), (i, W, ne)
        have = {t.tobytes() for t in tours}
        for j in range(W):
            if j != i:
                for t in pre[j][0][:ne]:
                    assert t.tobytes() in have                 # the migrant's tour itself arrived
        for t, f in zip(tours[:4], fit[:4]):
            assert close(f, orc.evaluate(t.astype(np.int64)))
        # hall of fame = best of (own best-ever, migrants)
        bt, bs = e.ga_best()
        assert close(bs, -orc.evaluate(bt.astype(np.int64)))
        assert -bs <= min(pre[j][1][0] for j in range(W)) + 1e-18
    # keep going: later exchanges keep every island sound
    sts = _advance_all(engs, 5 * K + 3)
    for i, e in enumerate(engs):
        assert sts[i].generations == 6 * K + 3 and e.ga_island_epochs() == 6 and e.ga_selfcheck() == 0
        e.ga_end(); e.close()


def test_run_islands_is_reproducible_and_stops_collectively(clm):
    import allhic_b200 as A
    orc = O.OracleCLM(FIX_CLM, FIX_IDS)
    init = np.random.default_rng(5).permutation(clm.N).astype(np.int32)
    engs = _engines(3, clm)
    seen = []
    prm = engs[0].ga_params(npop=64, ngen=150, seed=42, migrate_every=25, n_elite=2, log_every=100, phase=1)
    best, st = A.Engine.ga_run_islands(engs, init, prm, callback=lambda ph, g, s, t: seen.append((g, s, sorted(t.tolist()) == list(range(clm.N)))))
    assert sorted(best.tolist()) == list(range(clm.N))
    assert close(st.best_score, -orc.evaluate(best.astype(np.int64)))
    assert st.stopped == 1 and st.generations % 25 == 0          # the collective EarlyStop is taken at an exchange point
    assert [g for g, _, _ in seen] == list(range(0, st.generations + 1, 100)) and all(ok for *_, ok in seen)
    assert all(b[1] >= a[1] for a, b in zip(seen, seen[1:]))      # best over all islands never decreases
    best2, st2 = A.Engine.ga_run_islands(engs, init, prm)
    assert (best2 == best).all() and st2.best_score == st.best_score and st2.generations == st.generations
    # one island of the same size, same seed: the island run must not be worse on this easy group
    solo, st1 = engs[0].ga_run(init, engs[0].ga_params(npop=64, ngen=150, seed=42))
    assert st.best_score >= st1.best_score * 0.97
    assert st.evaluations > st1.evaluations * 1.5 or st.generations < st1.generations
    # max_gens ends all islands together even between exchange points
    _, st3 = A.Engine.ga_run_islands(engs, init, engs[0].ga_params(npop=64, ngen=5000, max_gens=61, seed=1, migrate_every=25, n_elite=1))
    assert st3.generations == 61 and st3.stopped == 1
    for e in engs:
        e.close()


def test_island_argument_errors(clm, monkeypatch):
    import allhic_b200 as A
    e = A.Engine(0); e.load(clm)
    init = np.arange(clm.N, dtype=np.int32)
    for kw in (dict(n_islands=2, island=2), dict(n_islands=2, island=0, migrate_every=0, n_elite=1),
               dict(n_islands=2, island=0, migrate_every=5, n_elite=0), dict(n_islands=33, island=0, migrate_every=5, n_elite=1),
               dict(n_islands=4, island=0, migrate_every=5, n_elite=10, npop=40)):
        with pytest.raises(A.AhxError) as ei:
            e.ga_begin(init, e.ga_params(**{"npop": 64, "migrate_every": 5, "n_elite": 1, **kw}))
        assert ei.value.code == -1
    # groups the persistent kernel cannot hold run two kernels per generation: no in-kernel exchange there
    monkeypatch.setenv("AHX_GA_RES", "0")
    with pytest.raises(A.AhxError) as ei:
        e.ga_begin(init, e.ga_params(npop=64, n_islands=2, island=0, migrate_every=5, n_elite=1))
    assert "persistent" in str(ei.value)
    e.close()


def test_islands_on_sub_tours(clm):
    """A resumed run optimises only the contigs its tour file lists (clm.go:400-417): islands on a sub-tour."""
    import allhic_b200 as A
    orc = O.OracleCLM(FIX_CLM, FIX_IDS)
    init = np.random.default_rng(8).permutation(clm.N)[:61].astype(np.int32)
    engs = _engines(2, clm)
    best, st = A.Engine.ga_run_islands(engs, init, engs[0].ga_params(npop=80, ngen=100, seed=3, migrate_every=10, n_elite=2))
    assert sorted(best.tolist()) == sorted(init.tolist())
    assert close(st.best_score, -orc.evaluate(best.astype(np.int64))) and st.best_score > -orc.evaluate(init.astype(np.int64))
    for e in engs:
        e.close()


def test_island_peer_that_never_arrives_times_out(clm, monkeypatch):
    import allhic_b200 as A
    monkeypatch.setenv("AHX_ISLAND_TIMEOUT_MS", "300")
    engs = _engines(2, clm)
    init = np.arange(clm.N, dtype=np.int32)
    for i, e in enumerate(engs):
        e.ga_begin(init, e.ga_params(npop=32, ngen=1 << 30, max_gens=1 << 40, island=i, n_islands=2, migrate_every=3, n_elite=1))
    A.Engine.ga_island_connect_local(engs)
    with pytest.raises(A.AhxError) as ei:
        engs[0].ga_advance(10)                                   # island 1 is never advanced
    assert ei.value.code == -3 and "timed out" in str(ei.value)
    for e in engs:
        e.ga_end(); e.close()


# ---------------------------------------------------------------- one process per GPU (CUDA IPC)
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _ipc_worker(rank, world, port, q):
    import torch.distributed as dist
    import allhic_b200 as A
    from allhic_b200.islands import IslandGA
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # host transport of the 64-byte handles only
    try:
        clm = A.CLM(FIX_CLM, FIX_IDS)
        eng = A.Engine(rank); eng.load(clm)
        init = np.random.default_rng(5).permutation(clm.N).astype(np.int32)
        out = []
        for run in range(2):                                       # second run: the mailbox mapping is reused
            ga = IslandGA(eng, init, eng.ga_params(npop=256, ngen=1 << 30, max_gens=1 << 40, seed=7 + run), migrate_every=10, n_elite=2)
            assert ga.mode == "device"
            tour, score, st = ga.run(120)
            bad = eng.ga_selfcheck()
            out.append((tour.tolist(), score, int(st.generations), ga.exchanges, bad))
            ga.close()
        q.put((rank, out))
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_islands_one_process_per_gpu_over_cuda_ipc(built_lib):
    import torch.multiprocessing as mp
    import allhic_b200 as A
    world = min(4, A._ffi.lib().ahx_device_count())
    if world < 2:
        pytest.skip("needs >= 2 GPUs: kernels of different processes do not run side by side on one GPU")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    orc = O.OracleCLM(FIX_CLM, FIX_IDS)
    for run in range(2):
        t0, s0, g0, x0, bad0 = res[0][run]
        for r in range(world):
            t, s, g, x, bad = res[r][run]
            assert (t, s, g, x) == (t0, s0, g0, x0) and bad == 0
        assert g0 == 120 and x0 == 12 and sorted(t0) == list(range(100))
        assert close(s0, -orc.evaluate(np.array(t0, np.int64)))
