"""CLM: the reference's parsed contact-map object (clm.go:32-41) on the host.

`CLM(clmfile, refile)` mirrors `NewCLM(Clmfile, REfile)` (clm.go:121); parsing
and the map semantics live in libahx (ahx_clm_parse, C++), this class only
exposes the frozen pair table.
"""
import ctypes as C

import numpy as np

from . import _ffi


class CLM:
    def __init__(self, clmfile=None, refile=None, nthreads=0, _handle=None):
        self._l = _ffi.lib()
        if _handle is not None:
            self.h = _handle
        else:
            h = C.c_void_p()
            _ffi.check(self._l.ahx_clm_parse(str(clmfile).encode(), str(refile).encode(), nthreads, C.byref(h)))
            self.h = h
        self.Clmfile, self.REfile = clmfile, refile
        self._cache = None

    @classmethod
    def parse_groups(cls, clmfile, refiles, nthreads=0):
        """One pass over a genome-wide CLM for all of its groups (ahx_clm_parse_groups): a list of CLM,
        each identical to CLM(clmfile, refiles[g])."""
        l = _ffi.lib()
        n = len(refiles)
        arr = (C.c_char_p * n)(*[str(r).encode() for r in refiles])
        out = (C.c_void_p * n)()
        _ffi.check(l.ahx_clm_parse_groups(str(clmfile).encode(), arr, n, nthreads, out))
        res = []
        for g in range(n):
            c = cls(_handle=C.c_void_p(out[g]))
            c.Clmfile, c.REfile = clmfile, refiles[g]
            res.append(c)
        return res

    @classmethod
    def from_sizes(cls, sizes):
        """Builder for callers with their own parser (ahx_clm_new + add_line)."""
        l = _ffi.lib()
        s = np.ascontiguousarray(sizes, np.int64)
        h = C.c_void_p()
        _ffi.check(l.ahx_clm_new(len(s), s.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(h)))
        return cls(_handle=h)

    def add_line(self, ai, bi, ao, bo, dists):
        d = np.ascontiguousarray(dists, np.int64)
        _ffi.check(self._l.ahx_clm_add_line(self.h, ai, bi, ord(ao), ord(bo), d.ctypes.data_as(C.POINTER(C.c_int64)), len(d)))
        self._cache = None

    def finalize(self):
        _ffi.check(self._l.ahx_clm_finalize(self.h))

    def close(self):
        if getattr(self, "h", None):
            self._l.ahx_clm_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Tigs
    @property
    def N(self):
        return self._l.ahx_clm_ntigs(self.h)

    def names(self):
        return [self._l.ahx_clm_tig_name(self.h, i).decode() for i in range(self.N)]

    def sizes(self):
        return np.array([self._l.ahx_clm_tig_size(self.h, i) for i in range(self.N)], np.int64)

    def tig_index(self, name):
        return self._l.ahx_clm_tig_index(self.h, name.encode())

    # ---- frozen maps
    def tables(self):
        """dict(pa, pb, nlinks, strand, slots[E,8], hist[H,12]) -- layout of include/ahx.h."""
        if self._cache is None:
            self.finalize()
            E = self._l.ahx_clm_npairs(self.h); H = self._l.ahx_clm_nhists(self.h)
            pa = np.zeros(E, np.int32); pb = np.zeros(E, np.int32); nl = np.zeros(E, np.int32)
            st = np.zeros(E, np.int8); slots = np.zeros((E, 8), np.int32); hist = np.zeros((H, 12), np.int32)
            p32 = C.POINTER(C.c_int32)
            _ffi.check(self._l.ahx_clm_get_pairs(self.h, pa.ctypes.data_as(p32), pb.ctypes.data_as(p32), nl.ctypes.data_as(p32),
                                                 st.ctypes.data_as(C.POINTER(C.c_int8)), slots.ctypes.data_as(p32)))
            _ffi.check(self._l.ahx_clm_get_hists(self.h, hist.ctypes.data_as(p32)))
            self._cache = dict(pa=pa, pb=pb, nlinks=nl, strand=st, slots=slots, hist=hist)
        return self._cache

    def M(self):
        """Dense contact matrix, clm.go:437-447 (host-side view for inspection)."""
        t = self.tables()
        M = np.zeros((self.N, self.N), np.int64)
        M[t["pa"], t["pb"]] = t["nlinks"]
        M[t["pb"], t["pa"]] = t["nlinks"]
        return M
