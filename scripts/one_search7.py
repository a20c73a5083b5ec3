#!/usr/bin/env python3
"""Profiling aid: one search_7lut on a synthetic state (args: n [mask depth])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _support as S
import sboxgates_b200 as sb
n = int(sys.argv[1]); depth = int(sys.argv[2]) if len(sys.argv) > 2 else 0
fixed = [(0, 1), (5, 0), (3, 1)][:depth]
eng = sb.LutEngine(0)
rng = sb.Xorshift1024(np.random.RandomState(1).bytes(128))
r = sb.search_7lut(eng, S.synthetic_state(n, seed=n), S.sbox_target(S.rijndael_sbox(), 0), S.mux_mask(fixed), [b for b, _ in fixed], rng)
print(r.found, r.tuples_feasible, eng.kernel_ms(1), eng.kernel_ms(3))
