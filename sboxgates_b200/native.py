"""ctypes binding of libsboxgates_b200.so (include/sboxgates_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C sboxgates_b200/csrc`.  If it is
missing, loading fails loudly -- there is deliberately no fallback implementation.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SBG_LIB: another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("SBG_LIB") or os.path.join(_HERE, "libsboxgates_b200.so")

SBG_OK = 0
SBG_LIST_CAP = 100000
SBG_KEY_NONE = (1 << 64) - 1


class NativeLibraryError(RuntimeError):
    pass


class SbgResult(C.Structure):
    _fields_ = [
        ("found", C.c_int32), ("ordering", C.c_int32), ("pos_outer", C.c_int32),
        ("pos_middle", C.c_int32), ("func_outer", C.c_uint8), ("func_middle", C.c_uint8),
        ("func_inner", C.c_uint8), ("inner_seen", C.c_uint8), ("gates", C.c_uint16 * 7),
        ("stale_outer", C.c_uint16), ("index", C.c_uint64), ("key", C.c_uint64),
        ("tuples_feasible", C.c_uint64), ("tuples_swept", C.c_uint64),
    ]


class SbgJob(C.Structure):
    _fields_ = [("slot", C.c_int32), ("flags", C.c_int32), ("order5", C.POINTER(C.c_uint8)),
                ("outer7", C.POINTER(C.c_uint8)), ("middle7", C.POINTER(C.c_uint8)),
                ("gate_order", C.POINTER(C.c_uint16))]


class SbgNodeResult(C.Structure):
    _fields_ = [("found_stage", C.c_int32), ("gates3", C.c_uint16 * 3), ("func3", C.c_uint8),
                ("seen3", C.c_uint8), ("key3", C.c_uint64), ("r5", SbgResult), ("r7", SbgResult)]


SBG_DO_SCAN3, SBG_DO_SEARCH5, SBG_DO_SEARCH7 = 1, 2, 4
SBG_LANES = 8

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)

# name -> (restype, argtypes): every symbol include/sboxgates_b200.h declares.
SIGNATURES = {
    "sbg_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "sbg_destroy": (None, [C.c_void_p]),
    "sbg_last_error": (C.c_char_p, [C.c_void_p]),
    "sbg_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sbg_launch_count": (C.c_uint64, [C.c_void_p]),
    "sbg_plan_tickets": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_uint64,
                                   C.POINTER(C.c_uint64)]),
    "sbg_weighted_tickets": (C.c_int, [C.c_int, C.c_uint32, C.POINTER(C.c_uint32)]),
    "sbg_last_kernel_ms": (C.c_float, [C.c_void_p, C.c_int]),
    "sbg_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "sbg_transfer_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "sbg_host_seconds": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "sbg_alu_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "sbg_search_node": (C.c_int, [C.c_void_p, C.POINTER(SbgJob), C.POINTER(SbgNodeResult)]),
    "sbg_search_batch": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(SbgJob),
                                   C.POINTER(SbgNodeResult)]),
    "sbg_list7_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]),
    "sbg_allgather_merge7": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int)]),
    "sbg_set_list7_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_int),
                                       C.c_int]),
    "sbg_load_problem": (C.c_int, [C.c_void_p, u64p, C.c_int, u64p, u64p, i8p]),
    "sbg_stage_problem": (C.c_int, [C.c_void_p, C.c_int, u64p, C.c_int, u64p, u64p, i8p]),
    "sbg_use_problem": (C.c_int, [C.c_void_p, C.c_int]),
    "sbg_search5": (C.c_int, [C.c_void_p, u8p, C.POINTER(SbgResult)]),
    "sbg_search7": (C.c_int, [C.c_void_p, u8p, u8p, C.POINTER(SbgResult)]),
    "sbg_search5_part": (C.c_int, [C.c_void_p, C.c_int, C.c_int, u8p, u64p]),
    "sbg_finish5": (C.c_int, [C.c_void_p, C.c_uint64, u8p, C.POINTER(SbgResult)]),
    "sbg_filter7_part": (C.c_int, [C.c_void_p, C.c_int, C.c_int, u64p, C.POINTER(C.c_int)]),
    "sbg_set_list7": (C.c_int, [C.c_void_p, u64p, C.c_int]),
    "sbg_decomp7_part": (C.c_int, [C.c_void_p, C.c_int, C.c_int, u8p, u8p, u64p]),
    "sbg_finish7": (C.c_int, [C.c_void_p, C.c_uint64, u8p, u8p, C.POINTER(SbgResult)]),
    "sbg_ordering_row": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "sbg_solve_inner": (C.c_int, [u64p, u64p, u64p, u64p, u64p, u8p, u8p]),
    "sbg_lut_table": (None, [C.c_uint8, u64p, u64p, u64p, u64p]),
}

_lib = None


def load_library(path=None):
    """Loads the shared library and binds every declared symbol; raises NativeLibraryError if the
    library or a symbol is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise NativeLibraryError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C sboxgates_b200/csrc`; sboxgates_b200 has no fallback implementation" % p)
    try:
        lib = C.CDLL(p)
    except OSError as exc:
        raise NativeLibraryError("cannot load %s: %s" % (p, exc)) from exc
    for name, (restype, argtypes) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise NativeLibraryError("%s lacks symbol %s" % (p, name)) from exc
        fn.restype = restype
        fn.argtypes = argtypes
    if path is None:
        _lib = lib
    return lib
