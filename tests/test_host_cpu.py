"""CPU: host-side logic and the C-ABI surface (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _support as S
import sboxgates_b200 as sb
from sboxgates_b200 import native


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(S.ROOT, "include", "sboxgates_b200.h")).read()
    declared = set(re.findall(r"\b(sbg_[a-z0-9_]+)\s*\(", header))
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    lib = native.load_library()          # raises if the .so or any symbol is missing
    for name in declared:
        assert getattr(lib, name) is not None


def test_result_struct_layout_matches_header():
    # int32 x4, uint8 x4, uint16 x7 + uint16, then four uint64: 72 bytes, no hidden padding.
    assert C.sizeof(native.SbgResult) == 72
    assert native.SbgResult.index.offset == 40


def test_ordering_rows_match_oracle():
    assert [sb.ordering_row(7, k) for k in range(70)] == S.order7_rows()
    assert [sb.ordering_row(5, k) for k in range(10)] == S.order5_rows()


def test_host_helpers_match_oracle():
    lib = S.oracle_lib()
    rs = np.random.RandomState(5)
    tabs = S.synthetic_state(16, seed=9)
    for i in range(200):
        a, b, c = (tabs[j] for j in rs.choice(16, 3, replace=False))
        f = int(rs.randint(0, 256))
        out = np.zeros(4, dtype=np.uint64)
        lib.orc_lut_ttable(f, S._u64(a)[1], S._u64(b)[1], S._u64(c)[1], out.ctypes.data_as(S.u64p))
        assert sb.lut_table(f, a, b, c).tolist() == out.tolist()
        mask = S.mux_mask([(int(rs.randint(0, 8)), int(rs.randint(0, 2)))]) if i % 2 else \
            np.full(4, np.uint64(2**64 - 1))
        tgt = S.lut_table(int(rs.randint(0, 256)), a, b, c) if i % 3 else tabs[int(rs.randint(16))]
        f2, s2 = C.c_uint8(), C.c_uint8()
        ok2 = lib.orc_solve_inner(S._u64(a)[1], S._u64(b)[1], S._u64(c)[1], S._u64(tgt)[1],
                                  S._u64(mask)[1], C.byref(f2), C.byref(s2))
        ok, func, seen = sb.solve_inner(a, b, c, tgt, mask)
        assert ok == bool(ok2)
        if ok:
            assert (func, seen) == (f2.value, s2.value)


def test_shuffles_consume_rng_like_reference():
    seed = open(os.path.join(S.GOLDEN, "seed1.bin"), "rb").read()
    r = sb.Xorshift1024(seed)
    order = sb.shuffled_order(r)
    assert r.draws == 256 and sorted(order) == list(range(256))
    o, m = sb.shuffled_orders7(r)
    assert r.draws == 256 + 512 and sorted(o) == list(range(256)) and sorted(m) == list(range(256))


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sb.NativeLibraryError):
        sb.LutEngine(0)


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(sb.NativeLibraryError):
        native.load_library(str(tmp_path / "nope.so"))


def test_ticket_plan_covers_the_allowed_prefixes_in_order():
    """sbg_plan_tickets (host planning code of the product, no device needed): the chunk tickets
    cover the lexicographically first prefixes made of allowed gates, and the whole-prefix tickets
    start at the rank -- among ALL prefixes -- of the first allowed prefix that is not covered."""
    import ctypes as C
    import itertools
    from math import comb
    lib = native.load_library()
    out = (C.c_uint64 * 4)()
    cases = [(7, 4, 30, 0b00101001, 1, 1), (7, 4, 30, 0, 1, 1), (7, 4, 26, 0b11, 1, 2),
             (7, 5, 24, 0b1001, 1, 1), (5, 3, 40, 0b01000101, 1, 1), (5, 3, 36, 0, 1, 1),
             (7, 4, 20, 0b101, 2, 1), (7, 4, 20, 0b101, 0, 1), (7, 4, 12, 0xff, 1, 1)]
    for width, pg, n, excluded, mode, waves in cases:
        assert lib.sbg_plan_tickets(width, pg, n, excluded, mode, waves, out) == 0
        items, chunks, t_offset, total = (int(x) for x in out)
        universe = n - (width - pg)                      # prefixes are pg-subsets of the first gates
        assert total == comb(universe, pg)
        if mode == 0:
            assert (items, t_offset) == (0, 0)
            continue
        allowed_gates = [g for g in range(universe) if not (g < 8 and (excluded >> g) & 1)]
        if n - bin(excluded & 0xff).count("1") < width:      # no combination of allowed gates at all
            assert items == 0
            continue
        qmax = comb(n - pg - 1, 2) if (width, pg) == (7, 4) else \
            (n - 6 if (width, pg) == (7, 5) else comb(n - 3, 2))
        assert chunks == max(1, (qmax + 31) // 32)
        assert items % chunks == 0
        covered = items // chunks
        n_allowed = comb(len(allowed_gates), pg)
        if n_allowed == 0:
            assert items == 0
            continue
        want_cover = n_allowed if mode == 2 else min(n_allowed, max(1, waves * 2368 // chunks))
        assert covered == want_cover
        # rank of the first uncovered allowed prefix among all prefixes (lexicographic order)
        if covered == n_allowed:
            assert t_offset == total
        else:
            first_left = next(itertools.islice(itertools.combinations(allowed_gates, pg), covered, None))
            rank = 0
            prev = -1
            for pos, g in enumerate(first_left):
                for y in range(prev + 1, g):
                    rank += comb(universe - y - 1, pg - pos - 1)
                prev = g
            assert t_offset == rank
            # everything below that rank is either covered or contains an excluded gate
            below = list(itertools.islice(itertools.combinations(range(universe), pg), rank))
            assert sum(1 for c in below if all(g in allowed_gates for g in c)) == covered
    assert lib.sbg_plan_tickets(6, 4, 30, 0, 1, 1, out) != 0


def test_weighted_ticket_tables_number_prefix_groups_in_order():
    """sbg_weighted_tickets (host planning code, no device needed): phase 1's weighted tickets cut a
    4-gate prefix (a,b,c,d) into ceil(C(n-d-2, 2) / group_pairs) groups and number (prefix, group) in
    lexicographic order.  The tables must unrank every ticket -- by the search the kernel does: per
    element, the largest y whose 'tickets before' count is <= t -- to exactly that enumeration."""
    import ctypes as C
    import itertools
    from math import comb
    lib = native.load_library()
    row = 76
    out = (C.c_uint32 * (1 + 4 * row))()
    for n, gp in ((7, 32), (9, 64), (16, 32), (22, 64), (30, 128), (40, 64)):
        assert lib.sbg_weighted_tickets(n, gp, out) == 0
        total = int(out[0])
        w = [[int(out[1 + row * r + x]) for x in range(row)] for r in range(4)]
        np_ = n - 3
        want = []
        for pre in itertools.combinations(range(np_), 4):
            pairs = comb(n - pre[3] - 2, 2)
            for g in range(max(1, -(-pairs // gp))):
                want.append((pre, g))
        assert total == len(want) == w[3][0]
        step = max(1, total // 4000)
        for t0 in list(range(0, total, step)) + [total - 1]:
            t, x0, pre = t0, 0, []
            for pos in range(4):
                r = 4 - pos
                wr = w[r - 1]
                e = max(y for y in range(x0, np_ - r + 1) if wr[x0] - wr[y] <= t)
                t -= wr[x0] - wr[e]
                pre.append(e)
                x0 = e + 1
            assert (tuple(pre), t) == want[t0], (n, gp, t0)
    assert lib.sbg_weighted_tickets(73, 32, out) != 0 and lib.sbg_weighted_tickets(20, 0, out) != 0


def test_graph_loader_and_verifier(tmp_path):
    """sboxgates_b200/graph.py: gates.xsd loader (the checks of state.c:260-411) + functional check."""
    from sboxgates_b200 import graph as G
    # 3-input "S-box" whose bit 0 is majority(a, b, c) and bit 1 is a ^ b
    sbox = [0] * 256
    for p in range(8):
        a, b, c = p & 1, (p >> 1) & 1, (p >> 2) & 1
        sbox[p] = (1 if a + b + c >= 2 else 0) | ((a ^ b) << 1)
    good = tmp_path / "g.xml"
    good.write_text(
        '<?xml version="1.0" encoding="UTF-8" ?>\n<gates>\n'
        '  <output bit="0" gate="3" />\n  <output bit="1" gate="4" />\n'
        '  <gate type="IN" />\n  <gate type="IN" />\n  <gate type="IN" />\n'
        '  <gate type="LUT" function="e8">\n    <input gate="2" />\n    <input gate="1" />\n'
        '    <input gate="0" />\n  </gate>\n'
        '  <gate type="XOR">\n    <input gate="0" />\n    <input gate="1" />\n  </gate>\n'
        '</gates>\n')
    g = G.load_graph(str(good))
    assert (g.num_inputs, g.num_luts, g.outputs) == (3, 1, {0: 3, 1: 4})
    assert G.verify_graph(g, sbox, require_bits=[0, 1]) == [0, 1]
    sbox_bad = list(sbox)
    sbox_bad[5] ^= 1
    import pytest
    with pytest.raises(G.GraphError):
        G.verify_graph(g, sbox_bad)
    bad = tmp_path / "b.xml"
    bad.write_text('<gates>\n  <gate type="IN" />\n  <gate type="AND">\n    <input gate="0" />\n'
                   '    <input gate="3" />\n  </gate>\n</gates>\n')
    with pytest.raises(G.GraphError):
        G.load_graph(str(bad))
