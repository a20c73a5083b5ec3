import os, sys
sys.path.insert(0, "/root/repo")
import bench, sboxgates_b200 as sb
eng = sb.LutEngine(0)
st = bench.build_batch(40, 8, 1003)[3]
eng.load(st["tables"], st["target"], st["mask"], st["inbits"])
for _ in range(2):
    r = eng.search7(st["outer"], st["middle"])
print("list", r.tuples_feasible, "found", r.found, hex(r.key))
