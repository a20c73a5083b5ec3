#!/usr/bin/env python3
"""Exploration aid (GPU): kernel times of full no-match sweeps for several n / masks."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _support as S
import sboxgates_b200 as sb
from math import comb

eng = sb.LutEngine(0)
sbox = S.rijndael_sbox()
tgt = S.sbox_target(sbox, 0)
for fixed in ([], [(0, 1)], [(0, 1), (5, 0)], [(0, 1), (5, 0), (3, 1)]):
    mask = S.mux_mask(fixed)
    inb = [b for b, _ in fixed]
    for n in [int(x) for x in sys.argv[1:]] or [24, 32, 48, 64, 96]:
        tabs = S.synthetic_state(n, seed=n)
        rng = sb.Xorshift1024(np.random.RandomState(1).bytes(128))
        t0 = time.time()
        r5 = sb.search_5lut(eng, tabs, tgt, mask, inb, rng)
        t1 = time.time()
        ms5 = eng.kernel_ms(0)
        r7 = sb.search_7lut(eng, tabs, tgt, mask, inb, rng)
        t2 = time.time()
        print("mask=%3d n=%3d | 5lut found=%d feas=%d wall=%.2fms k=%.3fms (%.2e T/s) | 7lut found=%d list=%d wall=%.2fms filter=%.3fms (%.2e T/s) sort=%.3f decomp=%.3fms" % (
            256 >> len(fixed), n, r5.found, r5.tuples_feasible, (t1 - t0) * 1e3, ms5, comb(n, 5) / max(ms5, 1e-6) * 1e3,
            r7.found, r7.tuples_feasible, (t2 - t1) * 1e3, eng.kernel_ms(1), comb(n, 7) / max(eng.kernel_ms(1), 1e-6) * 1e3,
            eng.kernel_ms(2), eng.kernel_ms(3)), flush=True)
