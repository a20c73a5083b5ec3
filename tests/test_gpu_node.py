"""GPU: the node-level entry points -- one device call chain per lut_search node (3-LUT scan +
search_5lut + search_7lut, lut.c:489-631), batches of independent nodes, device-resident gate
tables with incremental shipping, the segment path of very large sweeps, device-side list merge."""
import os
import subprocess
import sys

import numpy as np
import pytest

import _support as S
import sboxgates_b200 as sb
from sboxgates_b200.rng import Xorshift1024

pytestmark = pytest.mark.gpu

FULL = np.full(4, np.uint64(2**64 - 1), dtype=np.uint64)


def _cases(count, seed, n_choices=(7, 8, 9, 10, 11, 12, 13)):
    sbox = S.rijndael_sbox()
    rs = np.random.RandomState(seed)
    for i in range(count):
        n = int(rs.choice(n_choices))
        tabs = S.synthetic_state(n, seed=int(rs.randint(1 << 30)), num_inputs=min(8, n))
        fixed = [(int(b), int(rs.randint(0, 2))) for b in rs.choice(8, int(rs.randint(0, 4)),
                                                                    replace=False)]
        mask = S.mux_mask(fixed)
        inb = [b for b, _ in fixed if b < n]
        kind = int(rs.randint(0, 4))
        if kind == 0:   # a target some triple realises: the 3-LUT scan must find the FIRST such triple
            g = [int(x) for x in rs.choice(n, 3, replace=False)]
            tgt = S.lut_table(int(rs.randint(1, 255)), tabs[g[0]], tabs[g[1]], tabs[g[2]])
        elif kind == 1:  # a 5-input composition
            g = [int(x) for x in rs.choice(n, 5, replace=False)]
            tgt = S.lut_table(int(rs.randint(1, 255)),
                              S.lut_table(int(rs.randint(1, 255)), tabs[g[0]], tabs[g[1]], tabs[g[2]]),
                              tabs[g[3]], tabs[g[4]])
        else:
            tgt = S.sbox_target(sbox, int(rs.randint(0, 8)))
        order = [int(x) for x in rs.permutation(n)]
        yield tabs, tgt, mask, inb, order, rs.bytes(128)


def _reference_node(tabs, tgt, mask, inb, order, rng):
    """lut_search (lut.c:489-631) stage by stage with the CPU oracle: returns (stage, payload)."""
    n = len(tabs)
    for i in range(n):
        for k in range(i + 1, n):
            for m in range(k + 1, n):
                trip = [tabs[order[i]], tabs[order[k]], tabs[order[m]]]
                if S.oracle_check(3, tgt, mask, trip):
                    ok, func = S.oracle_get_lut_function(trip[0], trip[1], trip[2], tgt, mask, rng)
                    assert ok
                    return 3, (func, order[i], order[k], order[m])
    if n >= 5:
        found, ret, _ = S.oracle_search(5, tabs, tgt, mask, inb, rng)
        if found:
            return 5, ret
    if n >= 7:
        found, ret, _ = S.oracle_search(7, tabs, tgt, mask, inb, rng)
        if found:
            return 7, ret
    return 0, None


def test_node_call_matches_reference_stages(engine):
    """sbg_search_node == 3-LUT scan, then search_5lut, then search_7lut of the oracle, including
    the RNG draws each outcome consumes."""
    stages = {0: 0, 3: 0, 5: 0, 7: 0}
    for tabs, tgt, mask, inb, order, seed in _cases(60, 2024):
        o_rng = S.OrcRng.from_seed(seed)
        want_stage, want = _reference_node(tabs, tgt, mask, inb, order, o_rng)
        g_rng = Xorshift1024(seed)
        got = sb.lut_search(engine, tabs, tgt, mask, inb, order, g_rng)
        assert got.stage == want_stage, (got, want_stage, want)
        assert g_rng.draws == o_rng.draws
        if want_stage == 3:
            assert got.luts[0] == tuple(int(x) for x in want)
        elif want_stage == 5:
            assert [got.luts[0][0], got.luts[1][0]] + list(got.luts[0][1:]) + list(got.luts[1][2:]) \
                == want[:7]
        elif want_stage == 7:
            # luts = [middle, outer, inner] (the order the reference binary adds them in)
            assert [got.luts[1][0], got.luts[0][0], got.luts[2][0]] + list(got.luts[1][1:]) \
                + list(got.luts[0][1:]) + [got.luts[2][3]] == want
        stages[want_stage] += 1
    assert stages[3] > 5 and stages[5] > 5 and (stages[7] + stages[0]) > 5, stages


def test_batch_equals_single_calls(engine):
    """sbg_search_batch over staged states == the same searches one at a time."""
    sbox = S.rijndael_sbox()
    rs = np.random.RandomState(99)
    jobs, singles = [], []
    for slot in range(11):       # more jobs than lanes: two waves
        n = int(rs.choice([12, 20, 28, 33, 40]))
        tabs = S.synthetic_state(n, seed=300 + slot)
        fixed = [(int(b), int(rs.randint(0, 2))) for b in rs.choice(8, slot % 4, replace=False)]
        mask, inb = S.mux_mask(fixed), [b for b, _ in fixed]
        tgt = S.sbox_target(sbox, slot % 8)
        engine.stage(slot, tabs, tgt, mask, inb)
        o5 = bytes(rs.permutation(256).astype(np.uint8))
        oo = bytes(rs.permutation(256).astype(np.uint8))
        om = bytes(rs.permutation(256).astype(np.uint8))
        jobs.append(dict(slot=slot, order5=o5, outer=oo, middle=om))
    res = engine.search_batch(jobs)
    for j, r in zip(jobs, res):
        engine.use(j["slot"])
        r5 = engine.search5(j["order5"])
        assert (r.r5.found, r.r5.key, r.r5.tuples_feasible) == (r5.found, r5.key, r5.tuples_feasible)
        if not r5.found:
            r7 = engine.search7(j["outer"], j["middle"])
            assert (r.r7.found, r.r7.key, r.r7.tuples_feasible, list(r.r7.gates), r.r7.func_inner) \
                == (r7.found, r7.key, r7.tuples_feasible, list(r7.gates), r7.func_inner)
            assert r.r7.tuples_swept == r7.tuples_swept
        else:
            assert r.found_stage == 5 and not r.r7.found


def test_resident_tables_follow_a_growing_and_backtracking_state():
    """States of a graph build share a prefix of gates; only the difference is shipped (kernel
    arguments) and compression happens on the device.  A walk that appends gates, changes masks,
    backtracks (replaces a suffix) and jumps must give, at every step, what a fresh engine gives."""
    eng = sb.LutEngine(0)
    fresh = sb.LutEngine(0)
    sbox = S.rijndael_sbox()
    rs = np.random.RandomState(5)
    base = S.synthetic_state(70, seed=1)
    alt = S.synthetic_state(70, seed=2)
    n = 12
    tabs = base[:n].copy()
    before = eng.transfer_stats()
    steps = 0
    for step in range(40):
        move = int(rs.randint(0, 6))
        if move <= 2:                       # append 1-3 gates
            k = int(rs.randint(1, 4))
            src = base if rs.randint(0, 2) else alt
            tabs = np.concatenate([tabs, src[len(tabs):len(tabs) + k]])
        elif move == 3 and len(tabs) > 14:  # backtrack: drop a suffix, continue with other gates
            cut = int(rs.randint(10, len(tabs) - 2))
            tabs = np.concatenate([tabs[:cut], alt[cut:cut + 2]])
        elif move == 4:                     # jump (more than the argument space holds)
            tabs = S.synthetic_state(int(rs.randint(50, 64)), seed=100 + step)
        if len(tabs) > 66:
            tabs = tabs[:20].copy()
        fixed = [(int(b), int(rs.randint(0, 2))) for b in rs.choice(8, int(rs.randint(0, 4)),
                                                                    replace=False)]
        mask, inb = S.mux_mask(fixed), [b for b, _ in fixed]
        tgt = S.sbox_target(sbox, int(rs.randint(0, 8)))
        seed = rs.bytes(128)
        order = [int(x) for x in rs.permutation(len(tabs))]
        a = sb.lut_search(eng, tabs, tgt, mask, inb, order, Xorshift1024(seed))
        fresh.close()
        fresh = sb.LutEngine(0)
        b = sb.lut_search(fresh, tabs, tgt, mask, inb, order, Xorshift1024(seed))
        assert (a.stage, a.luts) == (b.stage, b.luts), step
        eng.load(tabs, tgt, mask, inb)
        fresh.load(tabs, tgt, mask, inb)
        assert np.array_equal(eng.filter7_part(0, 1), fresh.filter7_part(0, 1)), step
        steps += 1
    after = eng.transfer_stats()
    assert after[3] - before[3] > 5          # incremental shipments happened
    assert after[2] - before[2] < steps      # ... and bulk copies were the exception
    eng.close()
    fresh.close()


def test_sweeps_larger_than_the_ticket_table_run_in_segments():
    """A sweep with more tickets than the ticket table holds is cut into several launches, each
    appending to the list.  With a tiny table (SBG_TICKET_TABLE) the lists must equal the one-launch
    lists -- whole and in 3 parts, sparse and capped."""
    code = (
        "import sys, hashlib, json, numpy as np; sys.path[:0]=[%r, %r]\n"
        "import _support as S, sboxgates_b200 as sb\n"
        "eng = sb.LutEngine(0); sbox = S.rijndael_sbox(); out = []\n"
        "for n, fixed in [(34, [(0,1)]), (40, [(2,1),(3,0)]), (48, [(0,1),(5,0),(3,1)]), (70, [(1,1)])]:\n"
        "    eng.load(S.synthetic_state(n, seed=n), S.sbox_target(sbox, n %% 8), S.mux_mask(fixed), [b for b, _ in fixed])\n"
        "    whole = eng.filter7_part(0, 1)\n"
        "    parts = np.sort(np.concatenate([eng.filter7_part(p, 3) for p in range(3)]))[:100000]\n"
        "    r = eng.search7(bytes(range(256)), bytes(range(255, -1, -1)))\n"
        "    out.append([len(whole), hashlib.sha1(whole.tobytes()).hexdigest(), hashlib.sha1(parts.tobytes()).hexdigest(), int(r.key & 0xffffffffffff), int(r.tuples_feasible)])\n"
        "print(json.dumps(out))\n" % (S.ROOT, os.path.join(S.ROOT, "tests")))
    outs = []
    for table in ("", "4096"):
        env = dict(os.environ)
        if table:
            env["SBG_TICKET_TABLE"] = table
        res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                             check=True)
        outs.append(res.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1]
    import json
    rows = json.loads(outs[0])
    assert all(r[1] == r[2] for r in rows) and any(0 < r[0] < 100000 for r in rows)


def test_device_side_merge_of_part_lists(engine):
    """sbg_set_list7_device merges per-part lists without a host copy (what an all-gather into one
    device buffer gives): same installed list and same phase-2 key as the host-side path."""
    import torch
    sbox = S.rijndael_sbox()
    tabs = S.synthetic_state(30, seed=77)
    mask, inb = S.mux_mask([(1, 0), (4, 1)]), [1, 4]
    tgt = S.sbox_target(sbox, 3)
    engine.load(tabs, tgt, mask, inb)
    whole = engine.filter7_part(0, 1)
    parts = [engine.filter7_part(p, 4) for p in range(4)]
    stride = max(len(p) for p in parts)
    buf = torch.zeros((4, max(stride, 1)), dtype=torch.int64, device="cuda")
    for i, p in enumerate(parts):
        buf[i, :len(p)] = torch.from_numpy(p.view(np.int64)).cuda()
    torch.cuda.synchronize()
    outer, middle = sb.shuffled_orders7(Xorshift1024(np.random.RandomState(3).bytes(128)))
    engine.set_list7(np.concatenate(parts))
    k_host = engine.decomp7_part(0, 1, outer, middle)
    engine.set_list7_device(buf.data_ptr(), buf.shape[1], [len(p) for p in parts])
    ptr, cnt = engine.list7_device()
    assert cnt == len(whole)
    got = torch.empty(cnt, dtype=torch.int64, device="cuda")
    import ctypes
    ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(got.data_ptr()), ctypes.c_void_p(ptr),
                                           ctypes.c_size_t(8 * cnt), 3)
    assert np.array_equal(got.cpu().numpy().view(np.uint64), whole)
    assert engine.decomp7_part(0, 1, outer, middle) == k_host


def test_scan3_on_recorded_reference_states(engine):
    """Every recorded search_5lut call of a reference run comes from a lut_search whose 3-LUT scan
    had just failed (lut.c:501-523 precedes lut.c:553): the device scan must find nothing on those
    states either, whatever the gate order -- and, on the same states with a target that some triple
    does realise, it must return the first such triple of the given order (checked against the CPU
    oracle's check_n_lut_possible)."""
    rs = np.random.RandomState(7)
    recs = [r for r in S.read_records(os.path.join(S.GOLDEN, "run_rijndael_seed1.bin")) if r.which == 5]
    recs += [r for r in S.read_records(os.path.join(S.GOLDEN, "run_sodark_seed1.bin")) if r.which == 5]
    assert len(recs) > 100
    planted = 0
    for idx, rec in enumerate(recs):
        n = rec.n
        order = [int(x) for x in rs.permutation(n)]
        engine.load(rec.tables, rec.target, rec.mask, rec.inbits_list())
        node = engine.search_node(0, gate_order=order)
        assert node.found_stage == 0, (idx, n, list(node.gates3))
        if idx % 4 == 0:
            g = [int(x) for x in rs.choice(n, 3, replace=False)]
            tgt = S.lut_table(int(rs.randint(1, 255)), rec.tables[g[0]], rec.tables[g[1]],
                              rec.tables[g[2]])
            engine.load(rec.tables, tgt, rec.mask, rec.inbits_list())
            node = engine.search_node(0, gate_order=order)
            want = None
            for i in range(n):
                for k in range(i + 1, n):
                    for m in range(k + 1, n):
                        if want is None and S.oracle_check(3, tgt, rec.mask, [
                                rec.tables[order[i]], rec.tables[order[k]], rec.tables[order[m]]]):
                            want = (order[i], order[k], order[m])
                    if want:
                        break
                if want:
                    break
            assert want is not None and node.found_stage == 3
            assert tuple(int(x) for x in node.gates3) == want, (idx, n)
            planted += 1
    assert planted > 20
