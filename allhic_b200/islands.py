"""Island-model GA across GPUs (one process per GPU, torch.distributed).

The reference pins eaopt to a single population (`ga.NPops = 1`,
evaluate.go:269) although eaopt has multi-population + migration support;
SURVEY.md 8(e) maps that facility onto the 8 GPUs of a box: every rank runs
the device-resident GA of libahx on its own sub-population and Philox stream
(`island` = rank), and every `migrate_every` generations there is ONE exchange
step: an all-gather of each island's `n_elite` best tours (L int32 + one f64
each -- a few KB, latency-bound), after which every island replaces its worst
individuals by the other islands' elites.  The exchange is the only
collective; it runs over NCCL (NVLink/NVSwitch) when the process group is
NCCL and over gloo in the CPU tests.

Two exchange paths behind the same class:
  device   (product path) the exchange runs INSIDE libahx's persistent GA kernel: every
           island's kernel writes its elites into the other islands' mailboxes over NVLink
           peer memory and imports theirs, no host round trip and no relaunch (include/ahx.h
           "Islands").  The host only wires the mailboxes once per run: an all-gather of the
           64-byte CUDA IPC handles, which is also the barrier between consecutive runs.
  host     ga_get_elites -> all-gather -> ga_put_migrants between ga_advance calls; kept for
           groups the persistent kernel cannot hold (sub-tours, > 65535 contigs) and for the
           CPU protocol tests.

`engine` only needs ga_begin / ga_advance / ga_get_elites / ga_put_migrants /
ga_best / ga_end (plus ga_island_export / ga_island_connect for the device path), so the
protocol is testable on CPU with a stand-in engine.
"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

_TIMING = bool(os.environ.get("AHX_ISLAND_TIMING"))   # per-phase wall times of exchange() on stderr at close()


def _device_for(group):
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


class IslandGA:
    def __init__(self, engine, init_tour, params, migrate_every=50, n_elite=2, group=None):
        self.eng, self.group = engine, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.L = len(init_tour)
        self.migrate_every, self.n_elite = int(migrate_every), int(n_elite)
        params.island = self.rank                       # independent Philox stream per island
        self.params = params
        self.dev = _device_for(group)
        self.exchanges = 0
        self._t = [0.0] * 4
        init = np.ascontiguousarray(init_tour, np.int32)
        self.mode = "host"
        if self.world > 1 and hasattr(engine, "ga_island_export") and not os.environ.get("AHX_ISLANDS_HOST"):
            params.n_islands, params.migrate_every, params.n_elite = self.world, self.migrate_every, self.n_elite
            try:
                self.eng.ga_begin(init, params)
                self.mode = "device"
            except Exception as e:                      # group not on the persistent kernel: every rank falls back alike
                if "persistent" not in str(e):
                    raise
                params.n_islands = 0
        if self.mode == "host":
            if hasattr(params, "n_islands"):
                params.n_islands = 0
            self.eng.ga_begin(init, params)
        else:
            self._connect()

    def _connect(self):
        """All-gather the islands' mailbox handles and map them (once per run)."""
        mine = torch.from_numpy(np.ascontiguousarray(self.eng.ga_island_export(), np.uint8)).to(self.dev)
        allh = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        self.eng.ga_island_connect(allh.cpu().numpy().reshape(self.world, -1))

    def exchange(self, stopped=0):
        """All-gather elites (and this island's EarlyStop flag, so the stop decision costs no second
        collective); replace this island's worst by everyone else's best.  Returns (fitness of all
        elites [world, n_elite], True when every island has stopped)."""
        t0 = time.perf_counter()
        tours, fit = self.eng.ga_get_elites(self.n_elite)
        t1 = time.perf_counter()
        payload = torch.empty(self.n_elite, self.L + 3, dtype=torch.int32)
        payload[:, : self.L] = torch.from_numpy(tours)
        payload[:, self.L:self.L + 2] = torch.from_numpy(fit.view(np.int32).reshape(self.n_elite, 2))   # f64 bits as 2 x int32
        payload[:, self.L + 2] = int(stopped)
        payload = payload.to(self.dev)
        gathered = torch.empty(self.world * self.n_elite, self.L + 3, dtype=torch.int32, device=self.dev)
        dist.all_gather_into_tensor(gathered, payload, group=self.group)
        g = gathered.cpu().numpy().reshape(self.world, self.n_elite, self.L + 3)
        t2 = time.perf_counter()
        others = np.concatenate([g[r, :, : self.L] for r in range(self.world) if r != self.rank], axis=0) if self.world > 1 else None
        if others is not None and len(others):
            self.eng.ga_put_migrants(np.ascontiguousarray(others, np.int32))
        self.exchanges += 1
        t3 = time.perf_counter()
        self._t[0] += t1 - t0; self._t[1] += t2 - t1; self._t[2] += t3 - t2
        fits = np.ascontiguousarray(g[:, :, self.L:self.L + 2]).view(np.float64).reshape(self.world, self.n_elite)
        return fits, bool(np.all(g[:, 0, self.L + 2] != 0))

    def run(self, max_gens, on_exchange=None):
        """Advance all islands in lock step until every island's EarlyStop fired or
        `max_gens` generations ran.  Returns (best_tour, best_score, stats) with the
        global best replicated on every rank."""
        done, st = 0, None
        if self.mode == "device":
            t0 = time.perf_counter()
            st = self.eng.ga_advance(max_gens)           # exchanges and the collective EarlyStop happen inside the kernel
            self._t[3] += time.perf_counter() - t0
            self.exchanges = self.eng.ga_island_epochs() if hasattr(self.eng, "ga_island_epochs") else 0
            done = max_gens
        while done < max_gens:
            n = min(self.migrate_every, max_gens - done)
            t0 = time.perf_counter()
            st = self.eng.ga_advance(n)
            self._t[3] += time.perf_counter() - t0
            done += n
            if self.world > 1:
                fits, all_stopped = self.exchange(st.stopped)
                if on_exchange:
                    on_exchange(done, fits)
            else:
                all_stopped = bool(st.stopped)
            if all_stopped:
                break
        tour, score = self.eng.ga_best()
        best = torch.tensor([score], dtype=torch.float64, device=self.dev)
        allbest = torch.empty(self.world, dtype=torch.float64, device=self.dev)
        dist.all_gather_into_tensor(allbest, best, group=self.group)
        owner = int(torch.argmax(allbest).item())
        t = torch.from_numpy(np.ascontiguousarray(tour, np.int32)).to(self.dev)
        dist.broadcast(t, src=dist.get_global_rank(self.group, owner) if self.group is not None else owner, group=self.group)
        if self.world > 1 and self.mode == "host":
            st = self.eng.ga_advance(0)      # statistics after the last migrants
        return t.cpu().numpy(), float(allbest[owner].item()), st

    def close(self):
        if _TIMING and self.exchanges:
            import sys
            k = 1e3 / self.exchanges
            print("island %d: %d exchanges, ms each: ga_advance %.3f  get_elites %.3f  copy+all_gather+copy %.3f  put_migrants %.3f"
                  % (self.rank, self.exchanges, self._t[3] * k, self._t[0] * k, self._t[1] * k, self._t[2] * k), file=sys.stderr)
        self.eng.ga_end()


def farm_groups(n_groups, rank, world, weights=None):
    """One linkage group per GPU, no collective (allhic.go:170-173, 285-293):
    longest-processing-time-first assignment by weight (~N^2)."""
    order = sorted(range(n_groups), key=lambda g: -(weights[g] if weights is not None else 1))
    load = [0.0] * world
    mine = []
    for g in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += weights[g] if weights is not None else 1
        if r == rank:
            mine.append(g)
    return sorted(mine)
