"""ctypes loader for libahx.so (include/ahx.h).  No CPU fallback: importing
works without a GPU (so the ABI can be inspected), but creating an Engine
raises unless ahx_create finds a B200."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libahx.so")

OK = 0
ERR_ARG, ERR_CUDA, ERR_STATE, ERR_NOMEM, ERR_NODEVICE, ERR_IO = -1, -2, -3, -4, -5, -6
ISLAND_HANDLE_BYTES = 64


class AhxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libahx error %d: %s" % (code, msg))
        self.code = code


class GAParams(C.Structure):
    _fields_ = [("npop", C.c_int32), ("ngen", C.c_int32), ("max_gens", C.c_int64), ("mut_prob", C.c_double),
                ("seed", C.c_uint64), ("tournament_k", C.c_int32), ("log_every", C.c_int32), ("phase", C.c_int32),
                ("island", C.c_int32), ("n_islands", C.c_int32), ("migrate_every", C.c_int32), ("n_elite", C.c_int32),
                ("reserved", C.c_int32)]


class GAStats(C.Structure):
    _fields_ = [("best_score", C.c_double), ("generations", C.c_int64), ("updated", C.c_int64),
                ("evaluations", C.c_int64), ("stopped", C.c_int32), ("reserved", C.c_int32), ("device_ms", C.c_double)]


GA_CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.POINTER(C.c_int32), C.c_int32)

_p32 = C.POINTER(C.c_int32); _p64 = C.POINTER(C.c_int64); _pd = C.POINTER(C.c_double)
_pu8 = C.POINTER(C.c_uint8); _pu16 = C.POINTER(C.c_uint16); _pi8 = C.POINTER(C.c_int8); _vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/ahx.h declares
SIGNATURES = {
    "ahx_version": (C.c_int, []),
    "ahx_last_error": (C.c_char_p, [_vp]),
    "ahx_clm_parse": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(_vp)]),
    "ahx_clm_parse_groups": (C.c_int, [C.c_char_p, C.POINTER(C.c_char_p), C.c_int32, C.c_int, C.POINTER(_vp)]),
    "ahx_clm_new": (C.c_int, [C.c_int32, _p64, C.POINTER(_vp)]),
    "ahx_clm_add_line": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_uint8, C.c_uint8, _p64, C.c_int64]),
    "ahx_clm_finalize": (C.c_int, [_vp]),
    "ahx_clm_free": (None, [_vp]),
    "ahx_clm_ntigs": (C.c_int32, [_vp]),
    "ahx_clm_tig_name": (C.c_char_p, [_vp, C.c_int32]),
    "ahx_clm_tig_size": (C.c_int64, [_vp, C.c_int32]),
    "ahx_clm_tig_index": (C.c_int32, [_vp, C.c_char_p]),
    "ahx_clm_npairs": (C.c_int64, [_vp]),
    "ahx_clm_nhists": (C.c_int64, [_vp]),
    "ahx_clm_get_pairs": (C.c_int, [_vp, _p32, _p32, _p32, _pi8, _p32]),
    "ahx_clm_get_hists": (C.c_int, [_vp, _p32]),
    "ahx_device_count": (C.c_int, []),
    "ahx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "ahx_destroy": (None, [_vp]),
    "ahx_load": (C.c_int, [_vp, C.c_int32, _p64, C.c_int64, _p32, _p32, _p32, _pi8, _p32, C.c_int64, _p32]),
    "ahx_load_clm": (C.c_int, [_vp, _vp]),
    "ahx_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_uint64]),
    "ahx_host_free": (C.c_int, [_vp]),
    "ahx_eval_m": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, _vp]),
    "ahx_eval_m_u16": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, _vp]),
    "ahx_eval_m_dense": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, _vp]),
    "ahx_eval_q": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, C.c_int32, _vp, _vp]),
    "ahx_pop_set": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp]),
    "ahx_pop_eval_m": (C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
    "ahx_pop_scores": (C.c_int, [_vp, _vp]),
    "ahx_eval_m_kernel": (C.c_int, [_vp, C.c_int32, C.c_int32, _p32]),
    "ahx_pair_table_entries": (C.c_int64, [_vp, C.c_int32]),
    "ahx_flush_l2": (C.c_int, [_vp]),
    "ahx_ga_default_params": (None, [C.POINTER(GAParams)]),
    "ahx_ga_run": (C.c_int, [_vp, C.POINTER(GAParams), C.c_int32, _vp, _vp, C.POINTER(GAStats), GA_CALLBACK, _vp]),
    "ahx_ga_begin": (C.c_int, [_vp, C.POINTER(GAParams), C.c_int32, _vp]),
    "ahx_ga_advance": (C.c_int, [_vp, C.c_int64, C.POINTER(GAStats)]),
    "ahx_ga_best": (C.c_int, [_vp, _vp, _pd]),
    "ahx_ga_get_elites": (C.c_int, [_vp, C.c_int32, _vp, _vp]),
    "ahx_ga_put_migrants": (C.c_int, [_vp, C.c_int32, _vp]),
    "ahx_ga_island_export": (C.c_int, [_vp, _vp]),
    "ahx_ga_island_connect": (C.c_int, [_vp, _vp]),
    "ahx_ga_island_connect_local": (C.c_int, [C.POINTER(_vp), C.c_int32]),
    "ahx_ga_run_islands": (C.c_int, [C.POINTER(_vp), C.c_int32, C.POINTER(GAParams), C.c_int32, _vp, _vp, C.POINTER(GAStats), GA_CALLBACK, _vp]),
    "ahx_ga_island_epochs": (C.c_int64, [_vp]),
    "ahx_ga_selfcheck": (C.c_int, [_vp, _p64]),
    "ahx_ga_kernel": (C.c_int, [_vp]),
    "ahx_ga_end": (C.c_int, [_vp]),
    "ahx_mutate": (C.c_int, [_vp, C.c_int32, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp]),
    "ahx_mutate_remap": (C.c_int, [_vp, C.c_int32, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp]),
    "ahx_test_plan_mutations": (C.c_int, [C.c_uint64, C.c_int32, C.c_int64, C.c_int32, C.c_double, C.c_int32, _vp]),
    "ahx_test_tournaments": (C.c_int, [_vp, C.c_uint64, C.c_int32, C.c_int64, C.c_int32, _vp, C.c_int32, C.c_int32, _vp]),
    "ahx_flip_whole": (C.c_int, [_vp, C.c_int32, _vp, _vp, _pd, _pd, _p32]),
    "ahx_flip_one": (C.c_int, [_vp, C.c_int32, _vp, _vp, _pd, _p64, _p64, _vp, _vp, _p32]),
    "ahx_flip_all": (C.c_int, [_vp, C.c_int32, _vp, _vp, _pd, _pd, _p32]),
    "ahx_flip_all_info": (C.c_int, [_vp, _pd, _p32, _p32, _p32]),
    "ahx_probe_fp64": (C.c_int, [_vp, _pd, _pd]),
    "ahx_probe_l2": (C.c_int, [_vp, C.c_uint64, _pd]),
    "ahx_launch_count": (C.c_int64, [_vp]),
    "ahx_sm_count": (C.c_int, [_vp]),
    "ahx_last_device_ms": (C.c_double, [_vp]),
}

_lib = None


def lib():
    """Loads libahx.so; raises loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libahx.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(or `make -C allhic_b200/csrc`); there is no fallback path")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc, ctx=None):
    if rc != OK:
        msg = lib().ahx_last_error(ctx)
        raise AhxError(rc, msg.decode() if msg else "?")
