"""sboxgates_b200 -- B200-native 3-LUT exhaustive search (the `--lut` path of dansarie/sboxgates).

The product is `libsboxgates_b200.so` (hand-written sm_100a CUDA kernels behind the C ABI in
include/sboxgates_b200.h) plus `csrc/lut_shim.c`, which gives it the reference's own
`search_5lut` / `search_7lut` signatures (lut.h:46-55).  This package is the Python host-side mirror
of that interface, used by the tests, bench.py and the multi-GPU (one process per GPU) driver.
There is no CPU implementation in here: importing works without a GPU, constructing an engine
does not.
"""
from .rng import Xorshift1024
from .lut import LutEngine, SearchResult, NO_GATE, search_5lut, search_7lut, shuffled_order, \
    shuffled_orders7, ordering_row, solve_inner, lut_table, lut_search, LutSearchResult
from .native import load_library, NativeLibraryError

__all__ = [
    "Xorshift1024", "LutEngine", "SearchResult", "NO_GATE", "search_5lut", "search_7lut",
    "shuffled_order", "shuffled_orders7", "ordering_row", "solve_inner", "lut_table",
    "lut_search", "LutSearchResult",
    "load_library", "NativeLibraryError",
]
