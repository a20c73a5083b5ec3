import os, sys
sys.path.insert(0, "/root/repo")
import bench, sboxgates_b200 as sb
eng = sb.LutEngine(0)
for s in range(3, 6):
    for i, st in enumerate(bench.build_batch(40, 8, 1000 + s)):
        eng.load(st["tables"], st["target"], st["mask"], st["inbits"])
        r = eng.search7(st["outer"], st["middle"])
        print("depth", i % 4, "list", r.tuples_feasible, "found", r.found, flush=True)
