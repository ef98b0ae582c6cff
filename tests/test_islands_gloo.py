"""CPU, world_size 2 over gloo: the island-model exchange protocol
(allhic_b200/islands.py) and the group farm, with a stand-in engine.

The stand-in implements the six ahx_ga_* calls IslandGA needs (begin / advance /
get_elites / put_migrants / best / end) with a toy fitness, so what is tested
here is the host-side protocol: the payload packing (tour + f64 bits), the
all-gather, who replaces what, lock-step termination and the broadcast of the
global best.  The real engine behind the same protocol is covered by the
`-m gpu` tests (test_ga_elites_and_migrants) and by bench.py --gpus N.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from allhic_b200.islands import IslandGA, farm_groups


class _Stats:
    def __init__(self, best, gens, stopped, evals):
        self.best_score, self.generations, self.stopped, self.evaluations = best, gens, stopped, evals


class ToyEngine:
    """Score of a tour = -(number of positions i with tour[i] != i): identity is
    the optimum (0).  One 'generation' = P random swap proposals, keep improvements."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    @staticmethod
    def _score(t):
        return -float(np.count_nonzero(t != np.arange(len(t))))

    def ga_begin(self, init, prm):
        self.P, self.island = prm.npop, prm.island
        self.pop = np.repeat(np.asarray(init, np.int32)[None, :], self.P, 0).copy()
        self.fit = np.array([-self._score(t) for t in self.pop])      # fitness = -score, minimised
        self.gen = self.evals = 0
        self.best_t, self.best_f = self.pop[0].copy(), self.fit[0]
        self.max_gens = prm.max_gens

    def ga_advance(self, n):
        L = self.pop.shape[1]
        for _ in range(n):
            for p in range(self.P):
                i, j = self.rng.integers(0, L, 2)
                t = self.pop[p].copy(); t[i], t[j] = t[j], t[i]
                f = -self._score(t); self.evals += 1
                if f <= self.fit[p]:
                    self.pop[p], self.fit[p] = t, f
            self.gen += 1
            k = int(np.argmin(self.fit))
            if self.fit[k] < self.best_f:
                self.best_f, self.best_t = self.fit[k], self.pop[k].copy()
        return _Stats(-self.best_f, self.gen, int(self.best_f == 0.0 or self.gen >= self.max_gens), self.evals)

    def ga_get_elites(self, n):
        idx = np.argsort(self.fit, kind="stable")[:n]
        return self.pop[idx].copy(), self.fit[idx].copy()

    def ga_put_migrants(self, tours):
        L = self.pop.shape[1]
        for t in tours:
            assert sorted(t.tolist()) == list(range(L)), "migrant is not a permutation"
        idx = np.argsort(-self.fit, kind="stable")[: len(tours)]
        for k, t in zip(idx, tours):
            self.pop[k], self.fit[k] = t, -self._score(t)
        k = int(np.argmin(self.fit))
        if self.fit[k] < self.best_f:
            self.best_f, self.best_t = self.fit[k], self.pop[k].copy()

    def ga_best(self):
        return self.best_t.copy(), -self.best_f

    def ga_end(self):
        pass


class _Prm:
    npop, island, max_gens = 16, 0, 10_000


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = 24
        init = np.random.default_rng(5).permutation(L).astype(np.int32)      # same start on every island
        prm = _Prm()
        eng = ToyEngine(seed=100 + rank)
        seen = []
        ga = IslandGA(eng, init, prm, migrate_every=5, n_elite=2)
        assert prm.island == rank                                            # one RNG stream per island
        tour, score, st = ga.run(400, on_exchange=lambda gens, fits: seen.append((gens, fits.copy())))
        ga.close()
        # after an exchange every island holds the other island's elites
        other = 1 - rank
        g, fits = seen[0]
        assert g == 5 and fits.shape == (world, 2)
        assert np.all(fits[:, 0] <= fits[:, 1])                              # elites come sorted by fitness
        q.put((rank, tour.tolist(), score, ga.exchanges, float(fits[other, 0]), int(st.generations)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_island_exchange_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, t0, s0, x0, _, g0), (r1, t1, s1, x1, _, g1) = res
    assert (r0, r1) == (0, 1)
    assert t0 == t1 and s0 == s1, "the global best must be replicated on every rank"
    assert sorted(t0) == list(range(24)), "result must be a permutation (bit-exact validity)"
    assert x0 == x1 and x0 >= 1 and g0 == g1, "islands advance and exchange in lock step"
    assert s0 == -float(np.count_nonzero(np.array(t0) != np.arange(24)))     # reported score belongs to the reported tour


class ToyDeviceEngine(ToyEngine):
    """Stand-in for an engine whose exchange runs on the device (libahx's persistent GA kernel): IslandGA then
    only wires the islands -- export a 64-byte mailbox handle, all-gather, connect -- and advances them."""

    def ga_begin(self, init, prm):
        super().ga_begin(init, prm)
        assert prm.n_islands == dist.get_world_size() and prm.migrate_every == 5 and prm.n_elite == 2
        self.connected, self.epochs = None, 0

    def ga_island_export(self):
        h = np.zeros(64, np.uint8)
        h[:8] = np.frombuffer(np.int64(1000 + self.island).tobytes(), np.uint8)
        return h

    def ga_island_connect(self, handles):
        handles = np.asarray(handles)
        assert handles.shape == (dist.get_world_size(), 64) and handles.dtype == np.uint8
        self.connected = [int(np.frombuffer(handles[r, :8].tobytes(), np.int64)[0]) for r in range(handles.shape[0])]

    def ga_island_epochs(self):
        return self.epochs

    def ga_advance(self, n):
        assert self.connected == [1000 + r for r in range(dist.get_world_size())], "advance before the islands were wired"
        st = super().ga_advance(n)
        self.epochs = self.gen // 5
        return st


class _PrmDev(_Prm):
    n_islands, migrate_every, n_elite = 1, 50, 2


def _worker_device(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = 24
        init = np.random.default_rng(5).permutation(L).astype(np.int32)
        eng = ToyDeviceEngine(seed=200 + rank)
        ga = IslandGA(eng, init, _PrmDev(), migrate_every=5, n_elite=2)
        assert ga.mode == "device" and ga.params.island == rank and ga.params.n_islands == world
        tour, score, st = ga.run(60)
        ga.close()
        q.put((rank, tour.tolist(), score, ga.exchanges, int(st.generations)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_island_device_mode_wiring_world2_gloo():
    """The product path's host side: every island exports its mailbox handle, the handles are all-gathered in
    island order and connected before the first advance; one advance call runs the whole period (the exchange
    is the kernel's business); the global best is replicated on every rank."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_device, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, t0, s0, x0, g0), (r1, t1, s1, x1, g1) = res
    assert (r0, r1) == (0, 1) and t0 == t1 and s0 == s1 and sorted(t0) == list(range(24))
    assert x0 == x1 == 12 and g0 == g1 == 60
    assert s0 == -float(np.count_nonzero(np.array(t0) != np.arange(24)))


def test_farm_groups_partitions_every_group_once():
    w = [n * n for n in (700, 900, 850, 720, 810, 760, 880, 705, 799, 845, 731, 777)]
    for world in (1, 2, 3, 8):
        parts = [farm_groups(len(w), r, world, w) for r in range(world)]
        assert sorted(g for p in parts for g in p) == list(range(len(w)))
        loads = [sum(w[g] for g in p) for p in parts]
        assert max(loads) <= min(loads) + max(w)                             # LPT balance bound
    assert farm_groups(5, 0, 1) == [0, 1, 2, 3, 4]
