"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module; nothing under allhic_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_SO = os.path.join(ORACLE_DIR, "liboracle.so")


def build_oracle(force=False):
    src = os.path.join(ORACLE_DIR, "allhic_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "liboracle.so"])
    return _SO


class GAResult(C.Structure):
    _fields_ = [("best_score", C.c_double), ("generations", C.c_int64),
                ("evaluations", C.c_int64), ("seconds", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build_oracle())
    p64 = C.POINTER(C.c_int64); p32 = C.POINTER(C.c_int32); pd = C.POINTER(C.c_double); pu8 = C.POINTER(C.c_uint8)
    vp = C.c_void_p
    sig = {
        "orc_go_log": (C.c_double, [C.c_double]),
        "orc_sumlog": (C.c_double, [p64, C.c_int64]),
        "orc_golden_array": (None, [p64, C.c_int64, p64]),
        "orc_last_error": (C.c_char_p, []),
        "orc_new_clm": (vp, [C.c_char_p, C.c_char_p]),
        "orc_free_clm": (None, [vp]),
        "orc_ntigs": (C.c_int64, [vp]),
        "orc_tig_name": (C.c_char_p, [vp, C.c_int64]),
        "orc_tig_size": (C.c_int64, [vp, C.c_int64]),
        "orc_get_sizes": (None, [vp, p64]),
        "orc_ncontacts": (C.c_int64, [vp]),
        "orc_get_contacts": (None, [vp, p32, p32, p64, p64, pd]),
        "orc_noriented": (C.c_int64, [vp]),
        "orc_get_oriented": (None, [vp, p32, p32, pu8, pu8, p64]),
        "orc_set_active": (None, [vp, pu8]),
        "orc_set_signs": (None, [vp, pu8]),
        "orc_get_signs": (None, [vp, pu8]),
        "orc_tour_len": (C.c_int64, [vp]),
        "orc_get_tour": (None, [vp, p64]),
        "orc_get_M": (None, [vp, p64]),
        "orc_get_O": (None, [vp, pd]),
        "orc_set_tour": (None, [vp, C.c_int64, p64]),
        "orc_tour_evaluate": (C.c_double, [vp, C.c_int64, p64]),
        "orc_evaluate": (C.c_double, [vp]),
        "orc_count_W": (C.c_int64, [vp, C.c_int64, p64]),
        "orc_population_evaluate": (None, [vp, C.c_int64, C.c_int64, p64, pd, C.c_int]),
        "orc_mut_permute": (None, [p64, C.c_int64, C.c_int64, C.c_int64]),
        "orc_mut_splice": (None, [p64, C.c_int64, C.c_int64]),
        "orc_mut_insertion": (None, [p64, C.c_int64, C.c_int64, C.c_int64, C.c_int]),
        "orc_mut_inversion": (None, [p64, C.c_int64, C.c_int64, C.c_int64]),
        "orc_shuffle": (None, [p64, C.c_int64, C.c_uint64]),
        "orc_mutate_sample": (None, [C.c_int64, C.c_uint64, C.c_int64, p64]),
        "orc_tournament_sample": (None, [pd, C.c_int64, C.c_int, C.c_uint64, C.c_int64, p64]),
        "orc_ga_run": (C.c_int, [vp, C.c_int64, C.c_int64, C.c_double, C.c_uint64, C.c_int64, C.c_double, C.c_int, C.c_int, C.POINTER(GAResult)]),
        "orc_evaluate_q": (C.c_double, [vp]),
        "orc_evaluate_q_lookup": (C.c_double, [vp, p64]),
        "orc_flip_whole": (C.c_int, [vp, pd, pd]),
        "orc_flip_one": (C.c_int, [vp, p64, p64, pd, pd, pu8]),
        "orc_top_eigvec": (None, [vp, pd, pd]),
        "orc_flip_all": (C.c_int, [vp, pd, pd]),
        "orc_activate": (None, [vp, C.c_int, C.c_uint64, C.c_int]),
        "orc_format_tour": (C.c_int64, [vp, C.c_char_p, C.c_char_p, C.c_int64]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class OracleCLM:
    """The reference's CLM object (clm.go:32-41) restated on the CPU."""

    def __init__(self, clmfile, refile):
        self.l = lib()
        self.h = self.l.orc_new_clm(os.fsencode(clmfile), os.fsencode(refile))
        if not self.h:
            raise RuntimeError("oracle NewCLM failed: " + self.l.orc_last_error().decode())
        self.N = self.l.orc_ntigs(self.h)

    def close(self):
        if self.h:
            self.l.orc_free_clm(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- parsed state
    def names(self):
        return [self.l.orc_tig_name(self.h, i).decode() for i in range(self.N)]

    def sizes(self):
        out = np.zeros(self.N, np.int64)
        self.l.orc_get_sizes(self.h, _p(out, C.c_int64))
        return out

    def contacts(self):
        n = self.l.orc_ncontacts(self.h)
        ai = np.zeros(n, np.int32); bi = np.zeros(n, np.int32)
        st = np.zeros(n, np.int64); nl = np.zeros(n, np.int64); md = np.zeros(n, np.float64)
        self.l.orc_get_contacts(self.h, _p(ai, C.c_int32), _p(bi, C.c_int32), _p(st, C.c_int64), _p(nl, C.c_int64), _p(md, C.c_double))
        return ai, bi, st, nl, md

    def oriented(self):
        n = self.l.orc_noriented(self.h)
        ai = np.zeros(n, np.int32); bi = np.zeros(n, np.int32)
        ao = np.zeros(n, np.uint8); bo = np.zeros(n, np.uint8); g = np.zeros((n, 12), np.int64)
        self.l.orc_get_oriented(self.h, _p(ai, C.c_int32), _p(bi, C.c_int32), _p(ao, C.c_uint8), _p(bo, C.c_uint8), _p(g, C.c_int64))
        return ai, bi, ao, bo, g

    def M(self):
        out = np.zeros((self.N, self.N), np.int64)
        self.l.orc_get_M(self.h, _p(out, C.c_int64))
        return out

    def O(self):
        out = np.zeros((self.N, self.N), np.float64)
        self.l.orc_get_O(self.h, _p(out, C.c_double))
        return out

    # --- tour / signs
    def set_tour(self, idx):
        idx = np.ascontiguousarray(idx, np.int64)
        self.l.orc_set_tour(self.h, len(idx), _p(idx, C.c_int64))

    def tour(self):
        out = np.zeros(self.l.orc_tour_len(self.h), np.int64)
        self.l.orc_get_tour(self.h, _p(out, C.c_int64))
        return out

    def set_signs(self, signs):
        s = np.ascontiguousarray(signs, np.uint8)
        assert len(s) == self.N
        self.l.orc_set_signs(self.h, _p(s, C.c_uint8))

    def signs(self):
        out = np.zeros(self.N, np.uint8)
        self.l.orc_get_signs(self.h, _p(out, C.c_uint8))
        return out

    def set_active(self, active):
        a = np.ascontiguousarray(active, np.uint8)
        self.l.orc_set_active(self.h, _p(a, C.c_uint8))

    def activate(self, shuffle=True, seed=42, flip_all=True):
        self.l.orc_activate(self.h, int(shuffle), seed, int(flip_all))

    # --- scores
    def evaluate(self, idx=None):
        if idx is None:
            return self.l.orc_evaluate(self.h)
        idx = np.ascontiguousarray(idx, np.int64)
        return self.l.orc_tour_evaluate(self.h, len(idx), _p(idx, C.c_int64))

    def count_W(self, idx):
        idx = np.ascontiguousarray(idx, np.int64)
        return self.l.orc_count_W(self.h, len(idx), _p(idx, C.c_int64))

    def population_evaluate(self, tours, nthreads=1):
        t = np.ascontiguousarray(tours, np.int64)
        P, L = t.shape
        out = np.zeros(P, np.float64)
        self.l.orc_population_evaluate(self.h, P, L, _p(t, C.c_int64), _p(out, C.c_double), nthreads)
        return out

    def evaluate_q(self, literal=False):
        if literal:
            return self.l.orc_evaluate_q(self.h)
        return self.l.orc_evaluate_q_lookup(self.h, None)

    def evaluate_q_terms(self):
        n = C.c_int64(0)
        s = self.l.orc_evaluate_q_lookup(self.h, C.byref(n))
        return s, n.value

    def flip_whole(self):
        a = C.c_double(); b = C.c_double()
        acc = self.l.orc_flip_whole(self.h, C.byref(a), C.byref(b))
        return bool(acc), a.value, b.value

    def flip_one(self):
        L = self.l.orc_tour_len(self.h)
        na = C.c_int64(); nr = C.c_int64(); fs = C.c_double()
        trace = np.zeros((L, 2), np.float64); acc = np.zeros(L, np.uint8)
        any_ = self.l.orc_flip_one(self.h, C.byref(na), C.byref(nr), C.byref(fs), _p(trace, C.c_double), _p(acc, C.c_uint8))
        return bool(any_), na.value, nr.value, fs.value, trace, acc

    def flip_all(self):
        a = C.c_double(); b = C.c_double()
        acc = self.l.orc_flip_all(self.h, C.byref(a), C.byref(b))
        return bool(acc), a.value, b.value

    def top_eigvec(self):
        v = np.zeros(self.N, np.float64); w = C.c_double()
        self.l.orc_top_eigvec(self.h, _p(v, C.c_double), C.byref(w))
        return v, w.value

    def ga_run(self, npop=100, ngen=5000, mutprob=0.2, seed=42, max_gens=1000000, max_seconds=0.0, nthreads=1, verbose_phase=0):
        r = GAResult()
        rc = self.l.orc_ga_run(self.h, npop, ngen, mutprob, seed, max_gens, max_seconds, nthreads, verbose_phase, C.byref(r))
        if rc != 0:
            raise RuntimeError(self.l.orc_last_error().decode())
        return r

    def format_tour(self, label):
        cap = 64 + 64 * max(1, self.l.orc_tour_len(self.h))
        buf = C.create_string_buffer(cap)
        n = self.l.orc_format_tour(self.h, label.encode(), buf, cap)
        return buf.raw[:n].decode()


def golden_array(a):
    a = np.ascontiguousarray(a, np.int64)
    out = np.zeros(12, np.int64)
    lib().orc_golden_array(_p(a, C.c_int64), len(a), _p(out, C.c_int64))
    return out


def sumlog(a):
    a = np.ascontiguousarray(a, np.int64)
    return lib().orc_sumlog(_p(a, C.c_int64), len(a))


def go_log(x):
    return lib().orc_go_log(float(x))


def mutate(tour, op, p=0, q=0, k=0, front=0):
    t = np.ascontiguousarray(tour, np.int64).copy()
    L = len(t)
    l = lib()
    if op == "permute":
        l.orc_mut_permute(_p(t, C.c_int64), L, p, q)
    elif op == "splice":
        l.orc_mut_splice(_p(t, C.c_int64), L, k)
    elif op == "insertion":
        l.orc_mut_insertion(_p(t, C.c_int64), L, p, q, front)
    elif op == "inversion":
        l.orc_mut_inversion(_p(t, C.c_int64), L, p, q)
    else:
        raise ValueError(op)
    return t


def mutate_sample(L, seed, n):
    """n draws of Tour.Mutate's decisions (evaluate.go:221-232): [n, 4] = op, p, q, front."""
    out = np.zeros((n, 4), np.int64)
    lib().orc_mutate_sample(L, seed, n, _p(out, C.c_int64))
    return out


def tournament_sample(fit, k, seed, n):
    """n draws of eaopt SelTournament{k}.Apply(2): [n, 2] winners."""
    f = np.ascontiguousarray(fit, np.float64)
    out = np.zeros((n, 2), np.int64)
    lib().orc_tournament_sample(_p(f, C.c_double), len(f), k, seed, n, _p(out, C.c_int64))
    return out
