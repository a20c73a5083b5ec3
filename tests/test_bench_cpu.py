"""bench.py on the CPU side: the synthetic workload, the unit accounting and the reference arm's JSON
line (the GPU arm needs a device and is exercised by the driver)."""
import json
import math
import os
import subprocess
import sys
from types import SimpleNamespace

import numpy as np

import _support as S

sys.path.insert(0, S.ROOT)
import bench  # noqa: E402


def test_workload_is_deterministic_and_shaped_like_the_run():
    a = bench.build_batch(40, 8, 1000)
    b = bench.build_batch(40, 8, 1000)
    assert len(a) == 8
    for x, y in zip(a, b):
        assert np.array_equal(x["tables"], y["tables"]) and np.array_equal(x["mask"], y["mask"])
        assert x["order5"] == y["order5"] and x["outer"] == y["outer"] and x["middle"] == y["middle"]
    pops = sorted(int(sum(bin(int(w)).count("1") for w in st["mask"])) for st in a)
    assert pops == [32, 32, 64, 64, 128, 128, 256, 256]          # mux depth 3 .. 0, twice
    for st in a:
        assert st["tables"].shape == (40, 4)
        # the first 8 gates are the input bits, the target is S-box output bit 0
        assert np.array_equal(st["tables"][:8], S.synthetic_state(8, seed=1)[:8])
        assert np.array_equal(st["target"], S.sbox_target(S.rijndael_sbox(), 0))
        assert sorted(bytes(st["order5"])) == list(range(256))
        depth = {256: 0, 128: 1, 64: 2, 32: 3}[int(sum(bin(int(w)).count("1") for w in st["mask"]))]
        assert len([b for b in st["inbits"] if b >= 0]) == depth
    assert not np.array_equal(a[0]["tables"], bench.build_batch(40, 8, 1001)[0]["tables"])


def test_unit_accounting_follows_the_reference_enumeration():
    n = 20
    miss5 = SimpleNamespace(found=0, tuples_feasible=3, index=0, ordering=0, pos_outer=0)
    miss7 = SimpleNamespace(found=0, tuples_feasible=5, tuples_swept=math.comb(n, 7), index=0,
                            ordering=0, pos_outer=0, pos_middle=0)
    t, t7, c = bench.units_of(n, miss5, miss7)
    assert (t, t7) == (math.comb(n, 5), math.comb(n, 7))
    assert c == 3 * 10 * 256 + 5 * 70 * 65536
    hit5 = SimpleNamespace(found=1, tuples_feasible=9, index=41, ordering=2, pos_outer=7)
    hit7 = SimpleNamespace(found=1, tuples_feasible=5, tuples_swept=1234, index=3, ordering=4,
                           pos_outer=5, pos_middle=6)
    t, t7, c = bench.units_of(n, hit5, hit7)
    assert (t, t7) == (42, 1234)
    assert c == (2 * 256 + 7 + 1) + (3 * 70 * 65536 + 4 * 65536 + 5 * 256 + 6 + 1)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` on two host cores and a tiny state: one JSON line with the keys
    the driver reads (the default size is the driver's business, not this test's)."""
    code = (
        "import sys, json; sys.path.insert(0, %r); import bench\n"
        "r = bench.cpu_arm(14, 2, 5, budget_s=0.5, cores=2)\n"
        "print(json.dumps(r))\n" % S.ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True,
                         timeout=600).stdout.strip().splitlines()[-1]
    r = json.loads(out)
    assert r["unit"] == bench.UNIT and r["cores"] == 2 and r["kind"] in ("reference", "port")
    assert r["value"] > 0 and abs(r["value"] - (r["t_units_per_s"] + r["c_units_per_s"])) < 1e-6 * r["value"]
    assert "processes" in r["sample"]
