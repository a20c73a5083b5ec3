"""CPU: the multi-rank host logic (sboxgates_b200/distributed.py) over gloo, world_size 2 and 3,
with an oracle-backed stand-in for the CUDA engine.  What is tested here is the sharding / merge /
MIN-reduction / RNG lock-step logic, not the kernels (tests/test_gpu_parity.py does those)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import _support as S
from sboxgates_b200.distributed import DistributedLutSearch
from sboxgates_b200.lut import pack_tuple7, unpack_tuple7, SBG_KEY_NONE
from sboxgates_b200.native import SbgResult
from sboxgates_b200.rng import Xorshift1024


class OracleEngine:
    """Implements the LutEngine part-methods with the CPU oracle (tests only).  A part owns the
    combinations (5-LUT) / list entries (7-LUT) whose index is congruent to it modulo nparts."""

    def load(self, tables, target, mask, inbits):
        self.tables = np.ascontiguousarray(tables, dtype=np.uint64)
        self.target = np.ascontiguousarray(target, dtype=np.uint64)
        self.mask = np.ascontiguousarray(mask, dtype=np.uint64)
        self.inbits = list(inbits)
        self.n = self.tables.shape[0]
        self.list = None

    def search5_part(self, part, nparts, order):
        return S.oracle_search5_key(self.tables, self.target, self.mask, self.inbits, order, part,
                                    nparts)

    def finish5(self, key, order):
        import itertools
        res = SbgResult()
        res.key = key
        if key == SBG_KEY_NONE:
            return res
        n = self.tables.shape[0]
        comb = next(itertools.islice(itertools.combinations(range(n), 5), key >> 12, None))
        row = S.order5_rows()[(key >> 8) & 0xF]
        g = [comb[i] for i in row]
        res.found, res.ordering, res.pos_outer = 1, (key >> 8) & 0xF, key & 0xFF
        res.func_outer = order[key & 0xFF]
        t_outer = S.lut_table(res.func_outer, self.tables[g[0]], self.tables[g[1]], self.tables[g[2]])
        ok, f, seen = _solve(t_outer, self.tables[g[3]], self.tables[g[4]], self.target, self.mask)
        res.func_inner, res.inner_seen = f, seen
        for i in range(5):
            res.gates[i] = g[i]
        return res

    # -- 7-LUT ------------------------------------------------------------------------------------
    def filter7_part(self, part, nparts):
        lst, _ = S.oracle_filter7(self.tables, self.target, self.mask, self.inbits)
        mine = [pack_tuple7(t) for i, t in enumerate(lst.tolist()) if i % nparts == part]
        return np.array(mine, dtype=np.uint64)

    def filter7_keep_local(self):
        self.set_list7(self.filter7_part(0, 1))
        return len(self.list)

    def set_list7(self, packed):
        self.list = np.sort(np.asarray(packed, dtype=np.uint64))[:100000]

    def decomp7_part(self, part, nparts, outer, middle):
        tuples = np.array([unpack_tuple7(p) for p in self.list], dtype=np.uint16).reshape(-1, 7)
        return S.oracle_decomp7_key(self.tables, self.target, self.mask, tuples, outer, middle,
                                    part, nparts)

    def finish7(self, key, outer, middle):
        res = SbgResult()
        res.key = key
        res.tuples_feasible = len(self.list)
        if key == SBG_KEY_NONE:
            return res
        idx, k, po, pm = key >> 23, (key >> 16) & 0x7F, (key >> 8) & 0xFF, key & 0xFF
        t = unpack_tuple7(self.list[idx])
        g = [t[i] for i in S.order7_rows()[k]]
        tt = self.tables
        res.found, res.ordering, res.pos_outer, res.pos_middle = 1, k, po, pm
        res.func_outer, res.func_middle = outer[po], middle[pm]
        t_outer = S.lut_table(outer[po], tt[g[0]], tt[g[1]], tt[g[2]])
        t_mid = S.lut_table(middle[pm], tt[g[3]], tt[g[4]], tt[g[5]])
        ok, f, seen = _solve(t_outer, t_mid, tt[g[6]], self.target, self.mask)
        res.func_inner, res.inner_seen = f, seen
        for i in range(7):
            res.gates[i] = g[i]
        return res


def _solve(a, b, c, target, mask):
    import ctypes as C
    f, s = C.c_uint8(), C.c_uint8()
    ok = S.oracle_lib().orc_solve_inner(S._u64(a)[1], S._u64(b)[1], S._u64(c)[1],
                                        S._u64(target)[1], S._u64(mask)[1], C.byref(f), C.byref(s))
    return bool(ok), f.value, s.value


def _cases():
    """Small problems; sparse masks make matches appear, full ones make the sweeps complete."""
    out = []
    rs = np.random.RandomState(21)
    for i in range(5):
        n = 8 + i
        tabs = S.synthetic_state(n, seed=40 + i)
        mask = np.zeros(4, dtype=np.uint64)
        for p in rs.choice(256, [12, 16, 20, 40, 90][i], replace=False):
            mask[p >> 6] |= np.uint64(1) << np.uint64(p & 63)
        tgt = S.sbox_target(S.rijndael_sbox(), i)
        out.append((tabs, tgt, mask, [1] if i % 2 else []))
    return out


def _worker(rank, world, port, q, force_shard=True):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        results = []
        kw = dict(shard_min_tuples5=0, shard_min_tuples7=0, shard_min_list=0) if force_shard else {}
        drv = DistributedLutSearch(OracleEngine(), **kw)
        for tabs, tgt, mask, inb in _cases():
            for which in (5, 7):
                rng = Xorshift1024(np.random.RandomState(9).bytes(128))
                fn = drv.search_5lut if which == 5 else drv.search_7lut
                res = fn(tabs, tgt, mask, inb, rng)
                results.append((which, res.found, res.ret, rng.draws))
        q.put((rank, results, drv.collectives))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world,force_shard", [(2, True), (3, True), (2, False)])
def test_sharded_search_equals_single_rank_oracle(world, force_shard):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, force_shard))
             for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expected: the oracle's single-rank answer, same seed
    want = []
    for tabs, tgt, mask, inb in _cases():
        for which in (5, 7):
            rng = S.OrcRng.from_seed(np.random.RandomState(9).bytes(128))
            found, ret, st = S.oracle_search(which, tabs, tgt, mask, inb, rng)
            want.append((which, found, ret, int(rng.draws)))
    for rank, results, collectives in got:
        assert results == want, rank               # every rank holds the same, correct answer
        if force_shard:
            # 1 all-reduce per 5-LUT; per 7-LUT 1 count gather (+1 list gather if any hit) + 1 all-reduce
            assert len(_cases()) * 3 <= collectives <= len(_cases()) * 4
        else:
            assert collectives == 0   # small searches are replicated, not sharded
