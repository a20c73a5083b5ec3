import os, sys
sys.path.insert(0, "/root/repo")
import bench, sboxgates_b200 as sb
eng = sb.LutEngine(0)
for idx in (0, 3):
    st = bench.build_batch(40, 8, 1003)[idx]
    eng.load(st["tables"], st["target"], st["mask"], st["inbits"])
    for _ in range(2):
        r = eng.search5(st["order5"])
    print("depth", idx, "feasible", r.tuples_feasible, "found", r.found)
