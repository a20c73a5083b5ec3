#!/usr/bin/env python3
"""Large states under small masks (deep mux recursion): most combinations are feasible, the searches
end early.  Times search_5lut / search_7lut through the public API (GPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _support as S
import sboxgates_b200 as sb

eng = sb.LutEngine(0)
tgt = S.sbox_target(S.rijndael_sbox(), 0)
ns = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [64, 128, 200, 300, 500]
print("| n | mask | search_5lut ms (kernel) | found | search_7lut ms (phase 1 / sort / phase 2) | wall ms | list | found |")
print("|---|---|---|---|---|---|---|---|")
for n in ns:
    for fixed in ([(0, 1), (5, 0)], [(0, 1), (5, 0), (3, 1)], [(0, 1), (5, 0), (3, 1), (6, 1)],
                  [(0, 1), (5, 0), (3, 1), (6, 1), (2, 0)]):
        mask = S.mux_mask(fixed); inb = [b for b, _ in fixed]
        tabs = S.synthetic_state(n, seed=n)
        rng = sb.Xorshift1024(np.random.RandomState(1).bytes(128))
        r5 = sb.search_5lut(eng, tabs, tgt, mask, inb, rng); ms5 = eng.kernel_ms(0)
        t0 = time.perf_counter()
        r7 = sb.search_7lut(eng, tabs, tgt, mask, inb, rng)
        wall = 1e3 * (time.perf_counter() - t0)
        print("| %d | %d | %.3f | %d | %.3f / %.3f / %.3f | %.2f | %d | %d |" % (
            n, 256 >> len(fixed), ms5, r5.found, eng.kernel_ms(1), eng.kernel_ms(2), eng.kernel_ms(3),
            wall, r7.tuples_feasible, r7.found), flush=True)
