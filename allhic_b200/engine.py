"""Engine: one B200 + one loaded linkage group (wraps ahx_ctx, include/ahx.h)."""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import GAParams, GAStats, GA_CALLBACK


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


class PinnedArray:
    """numpy view of pinned host memory from ahx_host_alloc."""

    def __init__(self, shape, dtype):
        self._l = _ffi.lib()
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(shape)) * self.dtype.itemsize
        p = C.c_void_p()
        _ffi.check(self._l.ahx_host_alloc(C.byref(p), self.nbytes))
        self._p = p
        buf = (C.c_char * max(1, self.nbytes)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def close(self):
        if self._p:
            self.array = None
            self._l.ahx_host_free(self._p)
            self._p = None


class Engine:
    def __init__(self, device=0):
        self._l = _ffi.lib()
        h = C.c_void_p()
        _ffi.check(self._l.ahx_create(device, C.byref(h)))
        self.h = h
        self.N = 0
        self._cb_keepalive = None

    def close(self):
        if getattr(self, "h", None):
            self._l.ahx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _ffi.check(rc, self.h)

    # ---- upload
    def load(self, clm):
        clm.finalize()
        self._ck(self._l.ahx_load_clm(self.h, clm.h))
        self.N = clm.N

    def load_arrays(self, sizes, pa, pb, nlinks, strand, slots, hist):
        sizes = np.ascontiguousarray(sizes, np.int64); pa = np.ascontiguousarray(pa, np.int32); pb = np.ascontiguousarray(pb, np.int32)
        nlinks = np.ascontiguousarray(nlinks, np.int32); strand = np.ascontiguousarray(strand, np.int8)
        slots = np.ascontiguousarray(slots, np.int32); hist = np.ascontiguousarray(hist, np.int32)
        p32 = C.POINTER(C.c_int32)
        self._ck(self._l.ahx_load(self.h, len(sizes), sizes.ctypes.data_as(C.POINTER(C.c_int64)), len(pa), pa.ctypes.data_as(p32),
                                  pb.ctypes.data_as(p32), nlinks.ctypes.data_as(p32), strand.ctypes.data_as(C.POINTER(C.c_int8)),
                                  slots.ctypes.data_as(p32), len(hist), hist.ctypes.data_as(p32)))
        self.N = len(sizes)

    # ---- Tour.Evaluate, batched (evaluate.go:116)
    def eval_m(self, tours, out=None, kind="sparse"):
        t = np.ascontiguousarray(tours)
        if t.dtype == np.uint16 and kind == "sparse":
            fn = self._l.ahx_eval_m_u16
        else:
            t = np.ascontiguousarray(t, np.int32)
            fn = self._l.ahx_eval_m if kind == "sparse" else self._l.ahx_eval_m_dense
        if t.ndim == 1:
            t = t[None, :]
        P, L = t.shape
        if out is None:
            out = np.empty(P, np.float64)
        self._ck(fn(self.h, P, L, _ptr(t), _ptr(out)))
        return out

    # ---- CLM.EvaluateQ, batched (orientation.go:159)
    def eval_q(self, tours, signs):
        t = np.ascontiguousarray(tours, np.int32)
        s = np.ascontiguousarray(signs, np.uint8)
        if s.ndim == 1:
            s = s[None, :]
        P = s.shape[0]
        shared = 1 if t.ndim == 1 or t.shape[0] == 1 and P > 1 else 0
        if t.ndim == 1:
            t = t[None, :]
        assert s.shape[1] == self.N and (shared or t.shape[0] == P)
        out = np.empty(P, np.float64)
        self._ck(self._l.ahx_eval_q(self.h, P, t.shape[1], _ptr(t), shared, _ptr(s), _ptr(out)))
        return out

    # ---- resident population (kernel timing with inputs in HBM)
    def pop_set(self, tours):
        t = np.ascontiguousarray(tours, np.int32)
        self._ck(self._l.ahx_pop_set(self.h, t.shape[0], t.shape[1], _ptr(t)))
        self._pop_P = t.shape[0]

    def pop_eval_m(self, kind="sparse", repeat=1):
        ms = C.c_float()
        self._ck(self._l.ahx_pop_eval_m(self.h, 0 if kind == "sparse" else 1, repeat, C.byref(ms)))
        return ms.value

    def pop_scores(self):
        out = np.empty(self._pop_P, np.float64)
        self._ck(self._l.ahx_pop_scores(self.h, _ptr(out)))
        return out

    def eval_m_kernel(self, P, L=None):
        """(kind, tours_per_batch): 2 resident-table, 1 x32 streaming, 0 fp64 coordinates."""
        t = C.c_int32()
        kind = self._l.ahx_eval_m_kernel(self.h, P, self.N if L is None else L, C.byref(t))
        if kind < 0:
            self._ck(kind)
        return kind, t.value

    def pair_table_entries(self, tours_per_batch):
        """Entries (stored pairs + padding) of the resident kernel's pair table for T tours per batch."""
        return int(self._l.ahx_pair_table_entries(self.h, tours_per_batch))

    def flush_l2(self):
        self._ck(self._l.ahx_flush_l2(self.h))

    # ---- GA (evaluate.go:258)
    @staticmethod
    def ga_params(**kw):
        p = GAParams()
        _ffi.lib().ahx_ga_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def ga_run(self, init_tour, params, callback=None):
        t = np.ascontiguousarray(init_tour, np.int32)
        best = np.empty_like(t)
        st = GAStats()

        def _cb(user, phase, gen, score, tour_p, L):
            if callback:
                callback(phase, gen, score, np.ctypeslib.as_array(tour_p, shape=(L,)).copy())
        cb = GA_CALLBACK(_cb)
        self._ck(self._l.ahx_ga_run(self.h, C.byref(params), len(t), _ptr(t), _ptr(best), C.byref(st), cb, None))
        return best, st

    def ga_begin(self, init_tour, params):
        t = np.ascontiguousarray(init_tour, np.int32)
        self._ga_L = len(t)
        self._ck(self._l.ahx_ga_begin(self.h, C.byref(params), len(t), _ptr(t)))

    def ga_advance(self, n_gens):
        st = GAStats()
        self._ck(self._l.ahx_ga_advance(self.h, n_gens, C.byref(st)))
        return st

    def ga_best(self):
        t = np.empty(self._ga_L, np.int32)
        s = C.c_double()
        self._ck(self._l.ahx_ga_best(self.h, _ptr(t), C.byref(s)))
        return t, s.value

    def ga_get_elites(self, n):
        t = np.empty((n, self._ga_L), np.int32); f = np.empty(n, np.float64)
        self._ck(self._l.ahx_ga_get_elites(self.h, n, _ptr(t), _ptr(f)))
        return t, f

    def ga_put_migrants(self, tours):
        t = np.ascontiguousarray(tours, np.int32)
        self._ck(self._l.ahx_ga_put_migrants(self.h, t.shape[0], _ptr(t)))

    # ---- islands (include/ahx.h "Islands"): the exchange runs inside the GA kernel over peer memory
    def ga_island_export(self):
        """64-byte CUDA IPC handle of this island's mailbox (after ga_begin with n_islands > 1)."""
        h = np.zeros(_ffi.ISLAND_HANDLE_BYTES, np.uint8)
        self._ck(self._l.ahx_ga_island_export(self.h, _ptr(h)))
        return h

    def ga_island_connect(self, handles):
        """handles: [n_islands, 64] uint8, island order (all-gathered over any host transport)."""
        h = np.ascontiguousarray(handles, np.uint8)
        self._ck(self._l.ahx_ga_island_connect(self.h, _ptr(h)))

    def ga_island_epochs(self):
        return int(self._l.ahx_ga_island_epochs(self.h))

    @staticmethod
    def ga_island_connect_local(engines):
        """Islands of ONE process (one Engine per island, island order): peer access instead of CUDA IPC."""
        arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
        rc = _ffi.lib().ahx_ga_island_connect_local(arr, len(engines))
        for e in engines:
            if rc != _ffi.OK and _ffi.lib().ahx_last_error(e.h):
                _ffi.check(rc, e.h)
        _ffi.check(rc, engines[0].h)

    @staticmethod
    def ga_run_islands(engines, init_tour, params, callback=None):
        """CLM.GARun as an island model over the engines' GPUs, driven by this thread (ahx_ga_run_islands)."""
        t = np.ascontiguousarray(init_tour, np.int32)
        best = np.empty_like(t)
        st = GAStats()

        def _cb(user, phase, gen, score, tour_p, L):
            if callback:
                callback(phase, gen, score, np.ctypeslib.as_array(tour_p, shape=(L,)).copy())
        cb = GA_CALLBACK(_cb)
        arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
        _ffi.check(_ffi.lib().ahx_ga_run_islands(arr, len(engines), C.byref(params), len(t), _ptr(t), _ptr(best), C.byref(st), cb, None), engines[0].h)
        return best, st

    def ga_selfcheck(self):
        """Number of individuals of the current generation whose stored state is inconsistent (0 = sound)."""
        n = C.c_int64()
        self._ck(self._l.ahx_ga_selfcheck(self.h, C.byref(n)))
        return n.value

    def ga_kernel(self):
        """3/2 persistent kernel (pair table streamed / resident), 1 two kernels per generation, 0 fp64, -1 idle."""
        return self._l.ahx_ga_kernel(self.h)

    def ga_end(self):
        self._ck(self._l.ahx_ga_end(self.h))

    def mutate(self, tour, op, p=0, q=0, dir=0):
        t = np.ascontiguousarray(tour, np.int32)
        out = np.empty_like(t)
        self._ck(self._l.ahx_mutate(self.h, len(t), _ptr(t), op, p, q, dir, _ptr(out)))
        return out

    def mutate_remap(self, tour, op, p=0, q=0, dir=0):
        """The same mutation through the persistent GA kernel's operator code; returns (child tour, child 2*mid per contig)."""
        t = np.ascontiguousarray(tour, np.int32)
        out = np.empty_like(t); x = np.empty(self.N, np.uint32)
        self._ck(self._l.ahx_mutate_remap(self.h, len(t), _ptr(t), op, p, q, dir, _ptr(out), _ptr(x)))
        return out, x

    @staticmethod
    def plan_mutations(seed, island, gen, n, mut_prob, L):
        """[n, 5] = mutated, op, p, q, dir of the kernels' mutation plan (host-side evaluation of the same code)."""
        out = np.empty((n, 5), np.int32)
        _ffi.check(_ffi.lib().ahx_test_plan_mutations(seed, island, gen, n, mut_prob, L, _ptr(out)))
        return out

    def tournaments(self, fit, k, n_pairs, seed=1, island=0, gen=0):
        f = np.ascontiguousarray(fit, np.float64)
        w = np.empty((n_pairs, 2), np.int32)
        self._ck(self._l.ahx_test_tournaments(self.h, seed, island, gen, len(f), _ptr(f), k, n_pairs, _ptr(w)))
        return w

    # ---- orientation (orientation.go:31-120)
    def flip_whole(self, tour, signs):
        t = np.ascontiguousarray(tour, np.int32); s = np.ascontiguousarray(signs, np.uint8).copy()
        a = C.c_double(); b = C.c_double(); acc = C.c_int32()
        self._ck(self._l.ahx_flip_whole(self.h, len(t), _ptr(t), _ptr(s), C.byref(a), C.byref(b), C.byref(acc)))
        return s, bool(acc.value), a.value, b.value

    def flip_one(self, tour, signs):
        t = np.ascontiguousarray(tour, np.int32); s = np.ascontiguousarray(signs, np.uint8).copy()
        L = len(t)
        fs = C.c_double(); na = C.c_int64(); nr = C.c_int64(); acc = C.c_int32()
        trace = np.zeros((max(L, 1), 2), np.float64); steps = np.zeros(max(L, 1), np.uint8)
        self._ck(self._l.ahx_flip_one(self.h, L, _ptr(t), _ptr(s), C.byref(fs), C.byref(na), C.byref(nr), _ptr(trace), _ptr(steps), C.byref(acc)))
        return s, bool(acc.value), na.value, nr.value, fs.value, trace[:L], steps[:L]

    def flip_all(self, tour, signs):
        t = np.ascontiguousarray(tour, np.int32); s = np.ascontiguousarray(signs, np.uint8).copy()
        a = C.c_double(); b = C.c_double(); acc = C.c_int32()
        self._ck(self._l.ahx_flip_all(self.h, len(t), _ptr(t), _ptr(s), C.byref(a), C.byref(b), C.byref(acc)))
        return s, bool(acc.value), a.value, b.value

    def flip_all_info(self):
        """dict(residual, converged, dense, matvecs) of the last flip_all's eigenvector computation."""
        r = C.c_double(); c = C.c_int32(); d = C.c_int32(); m = C.c_int32()
        self._ck(self._l.ahx_flip_all_info(self.h, C.byref(r), C.byref(c), C.byref(d), C.byref(m)))
        return dict(residual=r.value, converged=bool(c.value), dense=bool(d.value), matvecs=m.value)

    # ---- probes
    def probe_fp64(self):
        a = C.c_double(); b = C.c_double()
        self._ck(self._l.ahx_probe_fp64(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def probe_l2(self, nbytes):
        g = C.c_double()
        self._ck(self._l.ahx_probe_l2(self.h, nbytes, C.byref(g)))
        return g.value

    def launch_count(self):
        return self._l.ahx_launch_count(self.h)

    def last_device_ms(self):
        """CUDA-event time of the kernels of the last eval_q / flip_one call."""
        return float(self._l.ahx_last_device_ms(self.h))

    def sm_count(self):
        return self._l.ahx_sm_count(self.h)
