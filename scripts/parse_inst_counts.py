#!/usr/bin/env python3
"""ncu CSV of the k_filter7_pm launches of one bench.py run -> per-state warp-instruction counts.

bench.py launches the kernel once per state, in a fixed order: the resident pass first (warm-up
steps, then the timed steps), each step `batch` states.  The table keeps the timed steps of that
first pass, keyed by the step's seed (1000 + warm-up + s), which is what bench.py looks up.
usage: parse_inst_counts.py filter_inst.csv steps warmup gates batch"""
import csv
import json
import sys

path, steps, warmup, gates, batch = sys.argv[1], *map(int, sys.argv[2:6])
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)
by_id = {}
for r in rows:
    k = int(r["ID"])
    by_id.setdefault(k, {"name": r["Kernel Name"]})
    val = float(r["Metric Value"].replace(",", ""))
    by_id[k][r["Metric Name"]] = val
launches = [by_id[k] for k in sorted(by_id)]
need = (warmup + steps) * batch
if len(launches) < need:
    sys.exit("only %d k_filter7_pm launches in %s, expected at least %d" % (len(launches), path, need))
per_step, alu, ns = {}, {}, {}
for s in range(steps):
    sl = launches[(warmup + s) * batch:(warmup + s + 1) * batch]
    per_step[str(1000 + warmup + s)] = [int(x["smsp__inst_executed.sum"]) for x in sl]
    alu[str(1000 + warmup + s)] = [int(x.get("smsp__inst_executed_pipe_alu.sum", 0)) for x in sl]
    ns[str(1000 + warmup + s)] = [int(x["gpu__time_duration.sum"]) for x in sl]
print(json.dumps({"gates": gates, "batch": batch, "steps": steps, "warmup": warmup,
                  "metric": "smsp__inst_executed.sum per k_filter7_pm launch (warp instructions)",
                  "per_step_seed": per_step, "alu_pipe_inst": alu, "ncu_duration_ns": ns,
                  "registers_per_thread": sorted({int(x.get("launch__registers_per_thread", 0))
                                                  for x in launches}),
                  "how": "scripts/ncu_inst_counts.sh"}, indent=1))
