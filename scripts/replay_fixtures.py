#!/usr/bin/env python3
"""Measurement (GPU): replays the recorded reference runs under tests/golden/ through the CUDA path
and sets the GPU wall time beside the reference's own time for the same calls (recorded per call by
oracle/_ref/sboxgates_rec on the build container's CPU, 1 rank).  Prints a markdown table."""
import glob, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import numpy as np
import _support as S
import sboxgates_b200 as sb
from sboxgates_b200 import lut as L
from sboxgates_b200.rng import Xorshift1024

eng = sb.LutEngine(0)
print("| fixture | calls (5-LUT / 7-LUT) | max n | reference, 1 core | GPU, through the C ABI | ratio | mismatches |")
print("|---|---|---|---|---|---|---|")
for path in sorted(glob.glob(os.path.join(S.GOLDEN, "run_*.bin")) + [os.path.join(S.GOLDEN, "ref_cases.bin")]):
    recs = S.read_records(path)
    if not recs:
        continue
    # pre-draw the orders (the caller's RNG work is not the library's)
    prepared = []
    for r in recs:
        rng = Xorshift1024.from_state(r.rng_s, r.rng_p)
        if r.which == 5:
            prepared.append((r, (L.shuffled_order(rng),), rng))
        else:
            prepared.append((r, L.shuffled_orders7(rng), rng))
    bad = 0
    for rep in range(2):           # second pass is the timed one (first warms caches / clocks)
        t0 = time.perf_counter()
        for r, orders, rng in prepared:
            eng.load(r.tables, r.target, r.mask, r.inbits_list())
            res = eng.search5(*orders) if r.which == 5 else eng.search7(*orders)
            if rep == 1 and bool(res.found) != r.found:
                bad += 1
            if rep == 1 and r.found:
                gates = [int(g) for g in res.gates[:r.which]]
                want = r.ret[2:7] if r.which == 5 else r.ret[3:10]
                if gates != want:
                    bad += 1
        gpu_s = time.perf_counter() - t0
    ref_s = sum(r.ns for r in recs) / 1e9
    n5 = sum(1 for r in recs if r.which == 5)
    print("| %s | %d / %d | %d | %.2f s | %.4f s | %.0fx | %d |" % (
        os.path.basename(path), n5, len(recs) - n5, max(r.n for r in recs), ref_s, gpu_s,
        ref_s / gpu_s, bad))
