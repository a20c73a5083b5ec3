"""Host-side mirror of the reference's LUT-search interface (lut.h:28-58) on top of the CUDA
library.

`search_5lut` / `search_7lut` take what the reference functions take -- the gate truth tables of
the state, target, mask, the used input bits -- plus the RNG the reference keeps as a static
(sboxgates.c:246-268), and return `(found, ret)` with `ret` the 10-entry array the reference fills
(lut.c:202-211, 453-462).  RNG consumption matches the reference call for call.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import native
from .native import (SbgResult, SbgJob, SbgNodeResult, NativeLibraryError, SBG_KEY_NONE,
                     SBG_LIST_CAP, SBG_DO_SCAN3, SBG_DO_SEARCH5, SBG_DO_SEARCH7)

NO_GATE = 0xFFFF  # state.h:30


@dataclass
class SearchResult:
    found: bool
    ret: List[int]                 # the reference's ret[10]
    ordering: int = -1
    key: int = SBG_KEY_NONE
    index: int = 0
    tuples_feasible: int = 0
    tuples_swept: int = 0
    stale_outer: bool = False
    gates: List[int] = field(default_factory=list)
    pos_outer: int = 0             # position of func_outer / func_middle in the shuffled orders
    pos_middle: int = 0


def shuffled_order(rng):
    """lut.c:125-135: Fisher-Yates over 0..255, one draw per element (256 draws)."""
    order = list(range(256))
    for i in range(256):
        j = rng.next() % (i + 1)
        order[i], order[j] = order[j], order[i]
    return bytes(order)


def shuffled_orders7(rng):
    """lut.c:362-378: outer and middle orders, draws interleaved (512 draws)."""
    outer = list(range(256))
    middle = list(range(256))
    for i in range(256):
        oj = rng.next() % (i + 1)
        mj = rng.next() % (i + 1)
        outer[i], outer[oj] = outer[oj], outer[i]
        middle[i], middle[mj] = middle[mj], middle[i]
    return bytes(outer), bytes(middle)


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(native.u64p)


def _order_ptr(order):
    buf = (C.c_uint8 * 256).from_buffer_copy(bytes(order))
    return buf


def ordering_row(width, k):
    lib = native.load_library()
    row = (C.c_int * width)()
    if lib.sbg_ordering_row(width, k, row) != 0:
        raise ValueError("bad ordering (%d, %d)" % (width, k))
    return [int(x) for x in row]


def lut_table(func, in1, in2, in3):
    """generate_lut_ttable (state.c:202-230)."""
    lib = native.load_library()
    _, p1 = _u64(in1)
    _, p2 = _u64(in2)
    _, p3 = _u64(in3)
    out = np.zeros(4, dtype=np.uint64)
    lib.sbg_lut_table(func, p1, p2, p3, out.ctypes.data_as(native.u64p))
    return out


def solve_inner(in1, in2, in3, target, mask):
    """get_lut_function without the random fill (lut.c:79-103): (ok, func, seen)."""
    lib = native.load_library()
    arrs = [_u64(x) for x in (in1, in2, in3, target, mask)]
    f = C.c_uint8()
    s = C.c_uint8()
    ok = lib.sbg_solve_inner(*[a[1] for a in arrs], C.byref(f), C.byref(s))
    return bool(ok), f.value, s.value


class LutEngine:
    """One CUDA device's search engine (an `sbg_handle`)."""

    def __init__(self, device=0, stream=None):
        self.lib = native.load_library()
        self._h = C.c_void_p()
        rc = self.lib.sbg_create(C.byref(self._h), int(device))
        if rc != 0:
            msg = self.lib.sbg_last_error(self._h).decode() if self._h else "sbg_create failed"
            if self._h:
                self.lib.sbg_destroy(self._h)
                self._h = C.c_void_p()
            raise NativeLibraryError("sbg_create(device=%d): %s" % (device, msg))
        self.device = device
        self.n = 0
        self._slot_n = {}
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sbg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("sboxgates_b200: %s (code %d)"
                               % (self.lib.sbg_last_error(self._h).decode(), rc))

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.sbg_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    @property
    def launches(self):
        return int(self.lib.sbg_launch_count(self._h))

    def kernel_ms(self, which):
        return float(self.lib.sbg_last_kernel_ms(self._h, which))

    def set_timing(self, on):
        """Kernel-family timing (CUDA events inside the chains; off by default)."""
        self._check(self.lib.sbg_set_timing(self._h, 1 if on else 0))

    def transfer_stats(self):
        """(h2d bytes, d2h bytes, bulk uploads, incremental uploads, unchanged states) so far."""
        out = (C.c_uint64 * 5)()
        self._check(self.lib.sbg_transfer_stats(self._h, out))
        return [int(x) for x in out]

    def alu_peak(self):
        """Measured LOP3 issue rate of the device, warp instructions per second."""
        v = C.c_double()
        self._check(self.lib.sbg_alu_peak(self._h, C.byref(v)))
        return float(v.value)

    # -- problem -------------------------------------------------------------------------------
    def load(self, tables, target, mask, inbits):
        tables, tp = _u64(tables)
        if tables.ndim != 2 or tables.shape[1] != 4:
            raise ValueError("tables must have shape (n, 4)")
        target, gp = _u64(target)
        mask, mp = _u64(mask)
        ib = np.full(8, -1, dtype=np.int8)
        ib[:len(inbits)] = inbits
        self.n = tables.shape[0]
        self._slot_n[0] = self.n
        self._check(self.lib.sbg_load_problem(self._h, tp, self.n, gp, mp,
                                              ib.ctypes.data_as(native.i8p)))

    def stage(self, slot, tables, target, mask, inbits):
        """Uploads a search state into device-resident slot `slot` without selecting it."""
        tables, tp = _u64(tables)
        target, gp = _u64(target)
        mask, mp = _u64(mask)
        ib = np.full(8, -1, dtype=np.int8)
        ib[:len(inbits)] = inbits
        self._check(self.lib.sbg_stage_problem(self._h, slot, tp, tables.shape[0], gp, mp,
                                               ib.ctypes.data_as(native.i8p)))
        self._slot_n[slot] = tables.shape[0]

    def prepare_state(self, tables, target, mask, inbits):
        """Marshals a state's host buffers once (numpy -> pointers) for stage_prepared: a caller that
        stages the same host arrays repeatedly, or wants the marshalling out of a timed region."""
        tables, tp = _u64(tables)
        target, gp = _u64(target)
        mask, mp = _u64(mask)
        ib = np.full(8, -1, dtype=np.int8)
        ib[:len(inbits)] = inbits
        return (tables, tp, target, gp, mask, mp, ib, ib.ctypes.data_as(native.i8p),
                int(tables.shape[0]))

    def stage_prepared(self, slot, prep):
        """stage() on the result of prepare_state (which keeps the host buffers alive)."""
        self._check(self.lib.sbg_stage_problem(self._h, slot, prep[1], prep[8], prep[3], prep[5],
                                               prep[7]))
        self._slot_n[slot] = prep[8]

    def use(self, slot):
        self._check(self.lib.sbg_use_problem(self._h, slot))
        self.n = self._slot_n[slot]

    # -- whole searches ------------------------------------------------------------------------
    def search5(self, func_order):
        res = SbgResult()
        self._check(self.lib.sbg_search5(self._h, _order_ptr(func_order), C.byref(res)))
        return res

    def search7(self, outer_order, middle_order):
        res = SbgResult()
        self._check(self.lib.sbg_search7(self._h, _order_ptr(outer_order),
                                         _order_ptr(middle_order), C.byref(res)))
        return res

    # -- one call per node / batches of nodes --------------------------------------------------
    @staticmethod
    def _job(slot, order5=None, outer=None, middle=None, gate_order=None):
        """Builds an SbgJob; returns (job, keepalive buffers)."""
        job = SbgJob()
        keep = []
        job.slot = slot
        flags = 0
        if gate_order is not None:
            go = (C.c_uint16 * len(gate_order))(*[int(g) for g in gate_order])
            keep.append(go)
            job.gate_order = C.cast(go, C.POINTER(C.c_uint16))
            flags |= SBG_DO_SCAN3
        if order5 is not None:
            b = _order_ptr(order5)
            keep.append(b)
            job.order5 = C.cast(b, C.POINTER(C.c_uint8))
            flags |= SBG_DO_SEARCH5
        if outer is not None:
            bo, bm = _order_ptr(outer), _order_ptr(middle)
            keep += [bo, bm]
            job.outer7 = C.cast(bo, C.POINTER(C.c_uint8))
            job.middle7 = C.cast(bm, C.POINTER(C.c_uint8))
            flags |= SBG_DO_SEARCH7
        job.flags = flags
        return job, keep

    def search_node(self, slot=0, order5=None, outer=None, middle=None, gate_order=None):
        """scan3 -> search_5lut -> search_7lut of one staged state as one device call chain."""
        job, keep = self._job(slot, order5, outer, middle, gate_order)
        res = SbgNodeResult()
        self._check(self.lib.sbg_search_node(self._h, C.byref(job), C.byref(res)))
        return res

    def search_batch(self, jobs):
        """jobs: list of dicts with keys slot, order5, outer, middle, gate_order (any may be absent).
        Returns the list of SbgNodeResult, one per job."""
        arr = (SbgJob * len(jobs))()
        keep = []
        for i, j in enumerate(jobs):
            job, k = self._job(j.get("slot", 0), j.get("order5"), j.get("outer"), j.get("middle"),
                               j.get("gate_order"))
            arr[i] = job
            keep.append(k)
        res = (SbgNodeResult * len(jobs))()
        self._check(self.lib.sbg_search_batch(self._h, len(jobs), arr, res))
        return list(res)

    def prepare_jobs(self, jobs):
        """The job array of search_batch, built once: (array, keep-alive buffers, count)."""
        arr = (SbgJob * len(jobs))()
        keep = []
        for i, j in enumerate(jobs):
            job, k = self._job(j.get("slot", 0), j.get("order5"), j.get("outer"), j.get("middle"),
                               j.get("gate_order"))
            arr[i] = job
            keep.append(k)
        return arr, keep, len(jobs)

    def search_batch_prepared(self, prepared):
        """search_batch on the result of prepare_jobs."""
        arr, _, count = prepared
        res = (SbgNodeResult * count)()
        self._check(self.lib.sbg_search_batch(self._h, count, arr, res))
        return list(res)

    def list7_device(self):
        """(device pointer, count) of this device's ordered phase-1 list."""
        ptr = C.c_void_p()
        cnt = C.c_int()
        self._check(self.lib.sbg_list7_device(self._h, C.byref(ptr), C.byref(cnt)))
        return ptr.value, cnt.value

    def set_list7_device(self, dev_ptr, stride, counts):
        """Merges ascending runs already in device memory (run r at dev_ptr + 8 * r * stride)."""
        arr = (C.c_int * len(counts))(*[int(c) for c in counts])
        self._check(self.lib.sbg_set_list7_device(self._h, C.c_void_p(dev_ptr), int(stride), arr,
                                                  len(counts)))

    # -- sharded building blocks ---------------------------------------------------------------
    def search5_part(self, part, nparts, func_order):
        key = C.c_uint64()
        self._check(self.lib.sbg_search5_part(self._h, part, nparts, _order_ptr(func_order),
                                              C.byref(key)))
        return key.value

    def finish5(self, key, func_order):
        res = SbgResult()
        self._check(self.lib.sbg_finish5(self._h, key, _order_ptr(func_order), C.byref(res)))
        return res

    def filter7_part(self, part, nparts):
        out = np.zeros(SBG_LIST_CAP, dtype=np.uint64)
        cnt = C.c_int()
        self._check(self.lib.sbg_filter7_part(self._h, part, nparts,
                                              out.ctypes.data_as(native.u64p), C.byref(cnt)))
        return out[:cnt.value].copy()

    def filter7_part_device(self, part, nparts):
        """Phase 1 of this part; the ordered list stays on the device.  Returns its length."""
        cnt = C.c_int()
        self._check(self.lib.sbg_filter7_part(self._h, part, nparts, None, C.byref(cnt)))
        return cnt.value

    def filter7_keep_local(self):
        """Phase 1 over the whole space on this device; the sorted, capped list stays in HBM as the
        installed list.  Returns its length."""
        cnt = C.c_int()
        self._check(self.lib.sbg_filter7_part(self._h, 0, 1, None, C.byref(cnt)))
        return cnt.value

    def set_list7(self, packed):
        packed, pp = _u64(packed)
        self._check(self.lib.sbg_set_list7(self._h, pp, int(packed.shape[0])))

    def decomp7_part(self, part, nparts, outer_order, middle_order):
        key = C.c_uint64()
        self._check(self.lib.sbg_decomp7_part(self._h, part, nparts, _order_ptr(outer_order),
                                              _order_ptr(middle_order), C.byref(key)))
        return key.value

    def finish7(self, key, outer_order, middle_order):
        res = SbgResult()
        self._check(self.lib.sbg_finish7(self._h, key, _order_ptr(outer_order),
                                         _order_ptr(middle_order), C.byref(res)))
        return res


def unpack_tuple7(packed):
    """63-bit packed 7-combination -> list of gate numbers."""
    p = int(packed)
    return [(p >> (9 * (6 - i))) & 0x1FF for i in range(7)]


def pack_tuple7(gates):
    p = 0
    for g in gates:
        p = (p << 9) | int(g)
    return p


def result5_to_ret(res, rng):
    """sbg_result -> the reference's ret[10] for search_5lut (lut.c:202-211), applying the random
    don't-care fill of get_lut_function (lut.c:104-106)."""
    if not res.found:
        return SearchResult(False, [0] * 10, key=int(res.key), tuples_feasible=int(res.tuples_feasible),
                            tuples_swept=int(res.tuples_swept))
    fi = res.func_inner
    if res.inner_seen != 0xFF:
        fi |= (~res.inner_seen & 0xFF) & (rng.next() & 0xFF)
    gates = [int(g) for g in res.gates[:5]]
    ret = [res.func_outer, fi] + gates + [0, 0, 0]
    return SearchResult(True, ret, ordering=res.ordering, key=int(res.key), index=int(res.index),
                        tuples_feasible=int(res.tuples_feasible),
                        tuples_swept=int(res.tuples_swept), gates=gates, pos_outer=int(res.pos_outer))


def result7_to_ret(res, rng):
    """sbg_result -> ret[10] for search_7lut (lut.c:453-462)."""
    if not res.found:
        return SearchResult(False, [0] * 10, key=int(res.key), tuples_feasible=int(res.tuples_feasible),
                            tuples_swept=int(res.tuples_swept))
    fi = res.func_inner
    if res.inner_seen != 0xFF:
        fi |= (~res.inner_seen & 0xFF) & (rng.next() & 0xFF)
    gates = [int(g) for g in res.gates[:7]]
    ret = [res.func_outer, res.func_middle, fi] + gates
    return SearchResult(True, ret, ordering=res.ordering, key=int(res.key), index=int(res.index),
                        tuples_feasible=int(res.tuples_feasible),
                        tuples_swept=int(res.tuples_swept), stale_outer=bool(res.stale_outer),
                        gates=gates, pos_outer=int(res.pos_outer), pos_middle=int(res.pos_middle))


def search_5lut(engine, tables, target, mask, inbits, rng):
    """lut.h:46-47.  Returns a SearchResult; `.found`, `.ret` are the reference's outputs."""
    if len(tables) < 5:
        raise ValueError("search_5lut needs at least 5 gates (lut.c:119)")
    order = shuffled_order(rng)
    engine.load(tables, target, mask, inbits)
    return result5_to_ret(engine.search5(order), rng)


def search_7lut(engine, tables, target, mask, inbits, rng):
    """lut.h:54-55."""
    if len(tables) < 7:
        raise ValueError("search_7lut needs at least 7 gates (lut.c:259)")
    # The reference draws its 512 shuffle values after phase 1 (lut.c:362-378); phase 1 itself
    # draws nothing, so drawing them first leaves the RNG stream unchanged.
    outer, middle = shuffled_orders7(rng)
    engine.load(tables, target, mask, inbits)
    return result7_to_ret(engine.search7(outer, middle), rng)


@dataclass
class LutSearchResult:
    """What lut_search() would add to the graph (lut.c:489-631): `luts` = the add_lut calls in
    order, each (function, in1, in2, in3) with inputs either gate numbers or ("new", k) = the k-th
    LUT added by this call; `stage` = 3, 5, 7 or 0 (NO_GATE)."""
    stage: int
    luts: List[tuple] = field(default_factory=list)
    node: object = None


def lut_search(engine, tables, target, mask, inbits, gate_order, rng, allow5=True, allow7=True):
    """lut.c:489-631 as ONE device call: the 3-LUT scan over the caller's shuffled gate order
    (lut.c:501-523), search_5lut (lut.c:553) and search_7lut (lut.c:593), each only if the earlier
    ones found nothing.  allow5 / allow7 = check_num_gates_possible(st, 2 / 3) (lut.c:525, 582).

    RNG: the reference draws 256 values on entry to search_5lut and 512 before phase 2 of
    search_7lut, plus one per solved LUT with unseen cells (lut.c:104-106).  Which of those happen
    depends on the stages' outcomes, so the shuffles are computed from a COPY of the generator
    (looking ahead) and the real one is advanced afterwards by exactly what the reference would
    have consumed."""
    n = len(tables)
    ahead = rng.copy()
    order5 = shuffled_order(ahead) if (allow5 and n >= 5) else None
    outer = middle = None
    if allow5 and allow7 and n >= 7:
        outer, middle = shuffled_orders7(ahead)
    engine.load(tables, target, mask, inbits)
    node = engine.search_node(0, order5, outer, middle, gate_order)
    if node.found_stage == 3:
        fi = node.func3
        if node.seen3 != 0xFF:
            fi |= (~node.seen3 & 0xFF) & (rng.next() & 0xFF)
        return LutSearchResult(3, [(fi, int(node.gates3[0]), int(node.gates3[1]),
                                    int(node.gates3[2]))], node)
    if order5 is None:
        return LutSearchResult(0, [], node)
    for _ in range(256):
        rng.next()
    if node.found_stage == 5:
        r = result5_to_ret(node.r5, rng).ret
        return LutSearchResult(5, [(r[0], r[2], r[3], r[4]), (r[1], ("new", 0), r[5], r[6])], node)
    if outer is None:
        return LutSearchResult(0, [], node)
    for _ in range(512):
        rng.next()
    if node.found_stage == 7:
        r = result7_to_ret(node.r7, rng).ret
        # lut.c:622-624 nests the outer and middle add_lut calls as arguments of the third; gcc
        # evaluates them right to left, so the MIDDLE LUT is added first (lower gate number)
        return LutSearchResult(7, [(r[1], r[6], r[7], r[8]), (r[0], r[3], r[4], r[5]),
                                   (r[2], ("new", 1), ("new", 0), r[9])], node)
    return LutSearchResult(0, [], node)
