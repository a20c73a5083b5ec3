#!/usr/bin/env python3
"""Measurement (GPU): SURVEY.md section 8d's synthetic sweeps.  Full no-match sweeps of search_5lut and
of search_7lut phase 1 for n in {32, 64, 96, 128} under mux masks of popcount 256/128/64/32, and the
phase-2 rate on the longest no-match list met.  Prints markdown tables."""
import os, sys
from math import comb
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _support as S
import sboxgates_b200 as sb

eng = sb.LutEngine(0)
tgt = S.sbox_target(S.rijndael_sbox(), 0)
rows, best_c = [], None
for n in (32, 64, 96, 128):
    for fixed in ([], [(0, 1)], [(0, 1), (5, 0)], [(0, 1), (5, 0), (3, 1)]):
        mask = S.mux_mask(fixed); inb = [b for b, _ in fixed]
        tabs = S.synthetic_state(n, seed=n)
        rng = sb.Xorshift1024(np.random.RandomState(1).bytes(128))
        r5 = sb.search_5lut(eng, tabs, tgt, mask, inb, rng); ms5 = eng.kernel_ms(0)
        r7 = sb.search_7lut(eng, tabs, tgt, mask, inb, rng)
        msf, msd = eng.kernel_ms(1), eng.kernel_ms(3)
        t5 = comb(n, 5) if not r5.found else r5.index + 1
        t7 = r7.tuples_swept
        rows.append((n, 256 >> len(fixed), t5, ms5, r5.found, t7, msf, r7.tuples_feasible, r7.found, msd))
        if not r7.found and r7.tuples_feasible > 1000 and (best_c is None or r7.tuples_feasible > best_c[0]):
            best_c = (r7.tuples_feasible, msd, n, 256 >> len(fixed))
print("| n | mask | 5-LUT tuples | 5-LUT ms | T5/s | 7-LUT tuples swept | phase-1 ms | T7/s | alg. GB/s (224 B/T) | list | found |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for n, m, t5, ms5, f5, t7, msf, lst, f7, msd in rows:
    print("| %d | %d | %.3g%s | %.3f | %.2e | %.3g | %.3f | %.2e | %.0f | %d | %d |" % (
        n, m, t5, "*" if f5 else "", ms5, t5 / ms5 * 1e3, t7, msf, t7 / msf * 1e3, t7 * 224 / msf * 1e3 / 1e9, lst, f7))
if best_c:
    lst, msd, n, m = best_c
    print("\nPhase 2 (no match, every candidate decided): %d listed tuples x 70 x 65,536 = %.3g C-units in %.3f ms "
          "= %.2e C-units/s (n = %d, mask %d)" % (lst, lst * 70 * 65536, msd, lst * 70 * 65536 / msd * 1e3, n, m))
