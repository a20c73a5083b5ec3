#!/usr/bin/env python3
"""How far into the lexicographic 7-combination space the 100,000th feasible tuple lies (GPU)."""
import os, sys
from math import comb
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _support as S
import sboxgates_b200 as sb

def rank7(t, n):
    r, prev = 0, -1
    for i, g in enumerate(t):
        for x in range(prev + 1, g):
            r += comb(n - x - 1, 6 - i)
        prev = g
    return r

eng = sb.LutEngine(0)
tgt = S.sbox_target(S.rijndael_sbox(), 0)
for n in (48, 64, 96, 128, 200):
    for fixed in ([(0, 1), (5, 0)], [(0, 1), (5, 0), (3, 1)], [(0, 1), (5, 0), (3, 1), (6, 1)]):
        mask = S.mux_mask(fixed); inb = [b for b, _ in fixed]
        tabs = S.synthetic_state(n, seed=n)
        eng.load(tabs, tgt, mask, inb)
        lst = eng.filter7_part(0, 1)
        ms = eng.kernel_ms(1)
        if len(lst) == 0:
            print(n, 256 >> len(fixed), "empty", ms); continue
        last = int(lst[-1])
        t = [(last >> (9 * (6 - i))) & 0x1ff for i in range(7)]
        print("n=%d m=%d list=%d last=%s rank=%.3g of %.3g  phase-1 %.3f ms swept %.3g" % (
            n, 256 >> len(fixed), len(lst), t, rank7(t, n), comb(n, 7), ms, eng.tuples_swept()
            if hasattr(eng, "tuples_swept") else -1))
