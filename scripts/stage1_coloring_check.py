#!/usr/bin/env python3
"""Design check for the next round (CPU only, no product code): the survivors of k_decomp7's stage 1
are exactly the proper 2-colourings of a conflict graph on the 8 outer patterns.

Stage 1 today: for an outer triple, pattern u of the three outer gates has 16-bit sets S1[u], S0[u]
over the patterns of the other four gates (cells holding a masked 1 / a masked 0).  An outer function
fo (the set X of patterns it maps to 1) survives iff neither X nor its complement merges a 1 and a 0:
    (OR_{u in X} S1[u]) & (OR_{u in X} S0[u]) == 0, and the same for ~X
-- evaluated for all 256 X by the subset sweep (172 instructions per (tuple, outer triple)).

Equivalent: u ~ u' iff S1[u] & S0[u'] or S0[u] & S1[u'] is non-empty; X is fine iff it is an
independent set; fo survives iff (X, ~X) is a proper 2-colouring.  So the survivors are empty if the
graph has an odd cycle, else 2^(number of connected components) functions, each a base colouring
XOR a union of component masks -- an adjacency build (28 pair tests), a breadth-first colouring of 8
vertices with bit masks, and the enumeration come out well under the subset sweep's cost.
This script checks the equivalence on random and on structured inputs."""
import itertools
import random


def survivors_subset_sweep(S1, S0):
    out = []
    for X in range(256):
        ok = True
        for side in (X, X ^ 0xFF):
            a1 = a0 = 0
            for u in range(8):
                if (side >> u) & 1:
                    a1 |= S1[u]
                    a0 |= S0[u]
            if a1 & a0:
                ok = False
        if ok:
            out.append(X)
    return out


def survivors_colouring(S1, S0):
    adj = [0] * 8
    for u, v in itertools.combinations(range(8), 2):
        if (S1[u] & S0[v]) or (S0[u] & S1[v]):
            adj[u] |= 1 << v
            adj[v] |= 1 << u
    colour1 = 0          # vertices coloured 1 in the base colouring (each component's root gets 0)
    seen = 0
    comps = []
    for root in range(8):
        if (seen >> root) & 1:
            continue
        comp, side = 1 << root, {root: 0}
        frontier = [root]
        while frontier:
            nxt = []
            for u in frontier:
                for v in range(8):
                    if (adj[u] >> v) & 1:
                        if v not in side:
                            side[v] = side[u] ^ 1
                            comp |= 1 << v
                            nxt.append(v)
                        elif side[v] == side[u]:
                            return []          # odd cycle: no outer function works
            frontier = nxt
        seen |= comp
        comps.append(comp)
        for v, s in side.items():
            if s:
                colour1 |= 1 << v
    out = []
    for pick in range(1 << len(comps)):
        X = colour1
        for j, comp in enumerate(comps):
            if (pick >> j) & 1:
                X ^= comp
        out.append(X)
    return sorted(out)


def main():
    rnd = random.Random(7)
    checked = nonempty = 0
    for trial in range(20000):
        density = rnd.choice([0.02, 0.05, 0.1, 0.2, 0.4])
        S1, S0 = [], []
        for u in range(8):
            a = sum(1 << b for b in range(16) if rnd.random() < density)
            z = sum(1 << b for b in range(16) if rnd.random() < density)
            z &= ~a                      # a cell never holds both (phase 1 guarantees feasibility)
            S1.append(a)
            S0.append(z)
        want = survivors_subset_sweep(S1, S0)
        got = survivors_colouring(S1, S0)
        assert want == got, (S1, S0, want, got)
        checked += 1
        nonempty += bool(want)
    print("stage-1 survivors == proper 2-colourings on %d random summaries (%d with survivors)"
          % (checked, nonempty))


if __name__ == "__main__":
    main()
