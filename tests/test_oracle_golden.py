"""CPU: the oracle (oracle/sbg_oracle.c) against the golden vectors produced by the reference
itself (oracle/gen_golden.py).  This is what pins the oracle."""
import glob
import json
import os

import numpy as np
import pytest

import _support as S

ORACLE_BUDGET_NS = 1.5e9  # replay only calls the reference itself answered this fast


def _records(pattern):
    out = []
    for path in sorted(glob.glob(os.path.join(S.GOLDEN, pattern))):
        for i, r in enumerate(S.read_records(path)):
            out.append((os.path.basename(path), i, r))
    return out


def _check(rec):
    rng = rec.rng()
    found, ret, st = S.oracle_search(rec.which, rec.tables, rec.target, rec.mask,
                                     rec.inbits_list(), rng)
    assert found == rec.found
    assert ret == rec.ret
    assert rng.draws == rec.draws  # RNG lock-step (lut.c:125-135, 362-378, 104-106)


def test_golden_files_present():
    names = {os.path.basename(p) for p in glob.glob(os.path.join(S.GOLDEN, "*"))}
    for need in ("ref_cases.bin", "primitives.json", "order_tables.json", "xml_names.json",
                 "run_crypto1_fc_seed1.bin", "run_des_s1_seed1.bin", "run_rijndael_seed1.bin"):
        assert need in names


def test_reference_cases():
    recs = _records("ref_cases.bin")
    assert len(recs) >= 60
    for name, i, rec in recs:
        _check(rec)


@pytest.mark.parametrize("pattern", ["run_crypto1_*.bin", "run_des_s1_seed1.bin",
                                     "run_rijndael_seed1.bin", "run_sodark_seed1.bin"])
def test_recorded_runs(pattern):
    recs = [x for x in _records(pattern) if x[2].ns < ORACLE_BUDGET_NS]
    assert recs
    for name, i, rec in recs:
        _check(rec)


def test_primitives():
    import ctypes as C
    lib = S.oracle_lib()
    vec = json.load(open(os.path.join(S.GOLDEN, "primitives.json")))
    for v in vec["lut_ttable"]:
        a, b, c = (np.array(x, dtype=np.uint64) for x in v["in"])
        out = np.zeros(4, dtype=np.uint64)
        lib.orc_lut_ttable(v["func"], S._u64(a)[1], S._u64(b)[1], S._u64(c)[1],
                           out.ctypes.data_as(S.u64p))
        assert out.tolist() == v["out"]
    for v in vec["get_lut_function"]:
        a, b, c = (np.array(x, dtype=np.uint64) for x in v["in"])
        t = np.array(v["target"], dtype=np.uint64)
        m = np.array(v["mask"], dtype=np.uint64)
        f = C.c_uint8()
        rng = S.OrcRng.from_seed(1)
        ok = lib.orc_get_lut_function(S._u64(a)[1], S._u64(b)[1], S._u64(c)[1], S._u64(t)[1],
                                      S._u64(m)[1], 0, C.byref(rng), C.byref(f))
        assert ok == v["ok"]
        if ok:
            assert f.value == v["func"]
        f2, s2 = C.c_uint8(), C.c_uint8()
        ok2 = lib.orc_solve_inner(S._u64(a)[1], S._u64(b)[1], S._u64(c)[1], S._u64(t)[1],
                                  S._u64(m)[1], C.byref(f2), C.byref(s2))
        assert ok2 == v["ok"]
        if ok2:
            assert f2.value == v["func"]  # randomize off: unseen cells stay 0 in both
    for v in vec["check_n_lut_possible"]:
        tabs = np.array(v["tables"], dtype=np.uint64)
        t = np.array(v["target"], dtype=np.uint64)
        m = np.array(v["mask"], dtype=np.uint64)
        assert lib.orc_check_n_lut_possible(v["num"], S._u64(t)[1], S._u64(m)[1],
                                            S._u64(tabs)[1]) == v["ok"]


def test_order_tables():
    literal = json.load(open(os.path.join(S.GOLDEN, "order_tables.json")))
    assert S.order7_rows() == literal                      # lut.c:396-415
    rows5 = S.order5_rows()
    assert rows5[0] == [0, 1, 2, 3, 4] and rows5[9] == [2, 3, 4, 0, 1]  # lut.c:189,224-229
    assert len({tuple(r[:3]) for r in rows5}) == 10


def test_rng_matches_reference_stream():
    # First outputs of xorshift1024* for the golden seed, cross-checked against the reference by the
    # recorded runs (their RNG states chain call to call); here: self-consistency of the two
    # implementations we ship (oracle C and sboxgates_b200.rng).
    from sboxgates_b200.rng import Xorshift1024
    seed = open(os.path.join(S.GOLDEN, "seed1.bin"), "rb").read()
    a = S.OrcRng.from_seed(seed)
    b = Xorshift1024(seed)
    lib = S.oracle_lib()
    for _ in range(1000):
        assert lib.orc_rng_next(a) == b.next()


def test_combination_rank_roundtrip():
    import ctypes as C
    lib = S.oracle_lib()
    for n, t in ((7, 7), (9, 5), (12, 7), (20, 5)):
        total = lib.orc_n_choose_k(n, t)
        prev = None
        for r in range(0, total, max(1, total // 200)):
            comb = (C.c_uint16 * t)()
            lib.orc_nth_combination(r, n, t, comb)
            assert lib.orc_combination_rank(n, t, comb) == r
            cur = list(comb)
            assert cur == sorted(cur) and len(set(cur)) == t
            if prev is not None:
                assert prev < cur
            prev = cur
