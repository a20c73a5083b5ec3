#!/usr/bin/env python3
"""bench.py -- candidate 3-LUT tuples/s of the `--lut` search path (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--gates n]

A STEP is one pass of the hot path -- search_5lut followed by search_7lut (lut.c:553,593) -- over a
batch of synthetic search states shaped like the ones `sboxgates --lut -o 0 rijndael.txt` presents
(BASELINE.json configs[1]): target = output bit 0 of the Rijndael S-box, n gates = 8 input bits +
random 3-LUTs of earlier gates (a real seeded run of that command ends with 31 LUTs, i.e. 39 gates,
profiles/r01_dropin_runs.md, hence the default n = 40), masks as mux recursion of depth 0..3 leaves
them (sboxgates.c:478,483), the selector bits excluded (lut.c:177-185).  No state has a 5-LUT
match, and at most a late 7-LUT one, so every step is a full sweep.

Units: T-unit = one gate combination through inbits-reject + feasibility;
C-unit = one candidate decomposition decided, counted as the reference enumerates them (2,560 per
feasible 5-tuple, 70 x 65,536 per listed 7-tuple) even though the kernels decide many at once from
a per-tuple summary.  value = (T + C) / s; both are also reported separately.

N > 1 (torchrun, one rank per GPU): the search states of a step are independent (in the program
they are the 16 mux branches of create_circuit, the 8 output bits of -o -1 and the -i iterations),
so every rank takes `--batch` states of a step of N x `--batch` and the ranks exchange only the
result keys (one all-gather per step) -- weak scaling.  `--shard tuples` instead keeps `--batch`
states in total and shards every single search over the ranks' GPUs (work items dealt round-robin;
one all-gather of the 7-LUT hit lists and one all-reduce(MIN) per search phase) -- strong scaling
of one search, which only pays for searches far larger than n = 40.

--impl reference times the reference's own object code (oracle/_ref/libsbgref.so, built from the
unmodified sources; the oracle port if that is absent) on the host cores, one process per core, each
on a bounded sample of the same workload: the same states restricted to their first n' gates.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidate 3-LUT tuples/sec on rijndael.txt --lut (T-units + C-units per second)"
UNIT = "tuples/s"
C_PER_5 = 10 * 256
C_PER_7 = 70 * 65536
BYTES_T5, BYTES_T7, BYTES_C = 160, 224, 160   # algorithmic bytes per unit


# ------------------------------------------------------------------------------------------------
# workload (self-contained: the product side must not import tests/ or oracle/)

def _input_table(bit):
    w = np.zeros(4, dtype=np.uint64)
    for p in range(256):
        if (p >> bit) & 1:
            w[p >> 6] |= np.uint64(1) << np.uint64(p & 63)
    return w


def _lut(func, a, b, c):
    full = np.uint64(0xFFFFFFFFFFFFFFFF)
    out = np.zeros(4, dtype=np.uint64)
    for m in range(8):
        if (func >> m) & 1:
            out |= (a if m & 4 else a ^ full) & (b if m & 2 else b ^ full) & (c if m & 1 else c ^ full)
    return out


def _rijndael_bit(bit):
    def mul(a, b):
        r = 0
        while b:
            if b & 1:
                r ^= a
            a <<= 1
            if a & 0x100:
                a ^= 0x11B
            b >>= 1
        return r
    inv = [0] * 256
    for a in range(1, 256):
        for b in range(1, 256):
            if mul(a, b) == 1:
                inv[a] = b
                break
    w = np.zeros(4, dtype=np.uint64)
    for p in range(256):
        x = inv[p]
        y = x
        for s in (1, 2, 3, 4):
            y ^= ((x << s) | (x >> (8 - s))) & 0xFF
        if ((y ^ 0x63) >> bit) & 1:
            w[p >> 6] |= np.uint64(1) << np.uint64(p & 63)
    return w


def _mux_mask(fixed):
    w = np.zeros(4, dtype=np.uint64)
    for p in range(256):
        if all(((p >> b) & 1) == v for b, v in fixed):
            w[p >> 6] |= np.uint64(1) << np.uint64(p & 63)
    return w


def _state(n, seed):
    rs = np.random.RandomState(seed)
    tabs = [_input_table(i) for i in range(8)]
    while len(tabs) < n:
        i, j, k = rs.choice(len(tabs), 3, replace=False)
        tabs.append(_lut(int(rs.randint(1, 255)), tabs[i], tabs[j], tabs[k]))
    return np.stack(tabs[:n]).astype(np.uint64)


def _orders(rs):
    return bytes(rs.permutation(256).astype(np.uint8)), bytes(rs.permutation(256).astype(np.uint8)), \
        bytes(rs.permutation(256).astype(np.uint8))


def build_batch(n, batch, step_seed):
    """`batch` search states; masks cycle through mux depth 0..3 (popcount 256, 128, 64, 32)."""
    target = _rijndael_bit(0)
    rs = np.random.RandomState(step_seed)
    out = []
    for i in range(batch):
        depth = i % 4
        bits = rs.choice(8, depth, replace=False)
        fixed = [(int(b), int(rs.randint(0, 2))) for b in bits]
        o5, oo, om = _orders(rs)
        out.append(dict(tables=_state(n, int(rs.randint(1 << 30))), target=target,
                        mask=_mux_mask(fixed), inbits=[b for b, _ in fixed], order5=o5, outer=oo,
                        middle=om))
    return out


def units_of(n, r5, r7):
    """(T, C) units of one state from the two results, as the reference would have enumerated."""
    t = math.comb(n, 5) if not r5.found else int(r5.index) + 1
    c = int(r5.tuples_feasible) * C_PER_5
    if r5.found:   # candidates of the matching tuple only (earlier feasible tuples are not counted)
        c = r5.ordering * 256 + r5.pos_outer + 1
    t7 = int(r7.tuples_swept)      # this rank's share; summed over ranks by the caller
    if r7.found:
        c += int(r7.index) * C_PER_7 + r7.ordering * 65536 + r7.pos_outer * 256 + r7.pos_middle + 1
    else:
        c += int(r7.tuples_feasible) * C_PER_7
    return t, t7, c


# ------------------------------------------------------------------------------------------------
# clocks

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4)
                          if r[3 + i].lower().startswith("active")})
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own object code on the host cores

def _cpu_worker(args):
    """One host core: the reference's search_5lut + search_7lut on the first n' gates of one
    state.  Returns (seconds, T-units, C-units)."""
    kind, tables, target, mask, inbits, seed = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _support as S
    rng = S.OrcRng.from_seed(seed)
    t_units = c_units = 0
    t0 = time.perf_counter()
    if kind == "reference":
        S.ref_lib()
        f5, r5, _ = S.ref_search(5, tables, target, mask, inbits, rng)
        f7, r7, _ = S.ref_search(7, tables, target, mask, inbits, rng)
    else:
        f5, r5, _ = S.oracle_search(5, tables, target, mask, inbits, rng)
        f7, r7, _ = S.oracle_search(7, tables, target, mask, inbits, rng)
    dt = time.perf_counter() - t0
    # Unit accounting (outside the timed region) with the oracle's counters: same semantics.
    _, _, s5 = S.oracle_search(5, tables, target, mask, inbits, S.OrcRng.from_seed(seed))
    lst, s7 = S.oracle_filter7(tables, target, mask, inbits)
    t_units = int(s5.tuples_filtered) + int(s7.tuples_filtered)
    c_units = int(s5.candidates)
    if f7:
        rng2 = S.OrcRng.from_seed(seed)
        S.oracle_search(5, tables, target, mask, inbits, rng2)
        _, _, s7b = S.oracle_search(7, tables, target, mask, inbits, rng2)
        c_units += int(s7b.candidates)
    else:
        c_units += len(lst) * C_PER_7
    return dt, t_units, c_units


def _pick_sample_gates(state, budget_s):
    """Largest n' whose reference run is expected to fit the budget: about 1e7 filter tuples/s and
    0.8 s per listed 7-tuple, using the oracle's fast phase-1 count."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _support as S
    best = 9
    for npr in range(9, min(30, state["tables"].shape[0]) + 1):  # C(30,7) = 2.0e6 filter tuples
        lst, _ = S.oracle_filter7(state["tables"][:npr], state["target"], state["mask"],
                                  state["inbits"])
        est = math.comb(npr, 7) / 1e7 + math.comb(npr, 5) / 5e6 + 0.8 * len(lst)
        if est > budget_s:
            break
        best = npr
    return best


def cpu_arm(n, batch, seed, budget_s=12.0, cores=None):
    """Times the reference (or the oracle port) on a bounded sample, one process per host core."""
    import multiprocessing as mp
    kind = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsbgref.so")) \
        else "port"
    cores = cores or os.cpu_count() or 1
    states = build_batch(n, max(batch, cores), seed)
    jobs = []
    npr_used = []
    for i in range(cores):
        st = states[i % len(states)]
        npr = _pick_sample_gates(st, budget_s)
        npr_used.append(npr)
        jobs.append((kind, st["tables"][:npr].copy(), st["target"], st["mask"],
                     [b for b in st["inbits"] if b < npr], 77 + i))
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = max(r[0] for r in res)
    t_units = sum(r[1] for r in res)
    c_units = sum(r[2] for r in res)
    return {"value": (t_units + c_units) / wall, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d processes, each the reference's search_5lut+search_7lut on the first "
                      "n'=%s gates of one workload state (n=%d); %.3g T-units + %.3g C-units in %.1f s"
                      % (cores, sorted(set(npr_used)), n, t_units, c_units, wall),
            "t_units_per_s": t_units / wall, "c_units_per_s": c_units / wall, "seconds": wall}


# ------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gates", type=int, default=40)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", default="states", choices=["states", "tuples"],
                    help="N > 1: states = every rank searches --batch independent states per step "
                         "(weak scaling); tuples = --batch states in total, every search sharded "
                         "over the tuple space (strong scaling)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    by_states = world > 1 and args.shard == "states"
    n_states = args.batch * world if by_states else args.batch
    scaling = "weak" if by_states or world == 1 else "strong"
    config = {"workload": "rijndael.txt --lut -o 0 shaped: %d states/step, n=%d gates, target = "
                          "S-box bit 0, mux masks of depth 0-3, full no-match sweeps of "
                          "search_5lut+search_7lut" % (n_states, args.gates),
              "gates": args.gates, "states_per_step": n_states, "states_per_gpu": args.batch,
              "parallelism": "1 GPU" if world == 1 else (
                  "%d ranks x %d independent search states per step "
                  "(one all-gather of the result keys per step)" % (world, args.batch) if by_states else
                  "%d ranks; every search sharded over the tuple space (all-gather of hit lists + "
                  "all-reduce(MIN) per phase above the size thresholds, replicated below)" % world),
              "l2": "a 256 MiB buffer is overwritten between steps (L2 flush); every step uses new states"}

    if args.impl == "reference":
        if rank != 0:
            return
        t0 = time.perf_counter()
        vals = []
        for s in range(max(1, min(args.steps, 2))):
            vals.append(cpu_arm(args.gates, args.batch, 1000 + s, budget_s=10.0))
        best = max(vals, key=lambda v: v["value"])
        line = {"impl": "reference", "metric": METRIC, "value": best["value"], "unit": UNIT,
                "n_gpus": args.gpus, "steps": len(vals), "warmup": 0,
                "ms_per_step": 1e3 * best["seconds"], "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "u32 bitwise", "data": "synthetic", "config": config,
                "cpu_baseline": best,
                "e2e": {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "seconds_total": time.perf_counter() - t0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import sboxgates_b200 as sb
    from sboxgates_b200.distributed import DistributedLutSearch

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = sb.LutEngine(local_rank)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    drv = DistributedLutSearch(eng) if world > 1 and not by_states else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    n, B = args.gates, n_states
    total_steps = args.warmup + args.steps
    batches = [build_batch(n, B, 1000 + s) for s in range(total_steps)]

    def run_step(states, resident, acc):
        """One step.  resident=True: states are already staged in HBM slots (value);
        False: host tables go through sbg_load_problem inside the step (e2e)."""
        keys = []
        for i, st in enumerate(states):
            if by_states and i // args.batch != rank:
                continue          # another rank's state
            if resident:
                eng.use(i % args.batch)
            else:
                eng.load(st["tables"], st["target"], st["mask"], st["inbits"])
            if drv is None:
                r5 = eng.search5(st["order5"])
                k5 = eng.kernel_ms(0)
                r7 = eng.search7(st["outer"], st["middle"])
            else:
                r5 = drv.search5_sharded(st["order5"])
                k5 = eng.kernel_ms(0)
                r7 = drv.search7_sharded(st["outer"], st["middle"])
            keys += [int(r5.key) & 0x7FFFFFFFFFFFFFFF, int(r7.key) & 0x7FFFFFFFFFFFFFFF]
            if acc is not None:
                t5, t7, c = units_of(n, r5, r7)
                acc["T"] += t5
                acc["C"] += c
                if drv is None or drv.last_phase1_sharded:
                    acc["T7"] += t7           # this rank's share (summed over ranks below)
                else:
                    acc["T7_rep"] += t7       # phase 1 replicated on every rank: count it once
                acc["ms5"] += k5
                acc["ms_filter"] += eng.kernel_ms(1)
                acc["ms_sort"] += eng.kernel_ms(2)
                acc["ms_decomp"] += eng.kernel_ms(3)
        if by_states:
            # every rank ends the step knowing every state's result, as the host program would
            per = 2 * ((len(states) + world - 1) // world)
            mine = torch.full((per,), -1, dtype=torch.int64, device="cuda")
            mine[:len(keys)] = torch.tensor(keys, dtype=torch.int64)
            allk = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allk, mine)

    def timed(resident):
        acc = {"T": 0, "C": 0, "T7": 0, "T7_rep": 0, "ms5": 0.0, "ms_filter": 0.0, "ms_sort": 0.0,
               "ms_decomp": 0.0}
        sampler = ClockSampler(local_rank)   # samples every 20 ms from the warm-up on
        sampler.start()
        for s in range(args.warmup):
            if resident:
                for i, st in enumerate(batches[s]):
                    if not by_states or i // args.batch == rank:
                        eng.stage(i % args.batch, st["tables"], st["target"], st["mask"], st["inbits"])
            run_step(batches[s], resident, None)
        launches0 = eng.launches
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(args.steps)]
        barrier()
        t_wall = 0.0
        for s in range(args.steps):
            states = batches[args.warmup + s]
            if resident:   # inputs resident in HBM before the timed region of this step starts
                for i, st in enumerate(states):
                    if not by_states or i // args.batch == rank:
                        eng.stage(i % args.batch, st["tables"], st["target"], st["mask"], st["inbits"])
            flush.fill_(s & 0xFF)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ev[s][0].record(stream)
            run_step(states, resident, acc)
            ev[s][1].record(stream)
            torch.cuda.synchronize()
            t_wall += time.perf_counter() - t0
        clocks = sampler.stop()
        barrier()
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        # every search ends with a device->host read of its result, so device time ~ wall time;
        # report the larger (it includes host-side launch gaps) and take the max over ranks
        ms = max(dev_ms, 1e3 * t_wall)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        acc["T7_local"], acc["ms_filter_local"] = acc["T7"] + acc["T7_rep"], acc["ms_filter"]
        if world > 1:   # units are produced on different ranks: sum the shares
            t = torch.tensor([acc["T7"]] + ([acc["T"], acc["C"]] if by_states else [0, 0]),
                             dtype=torch.int64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            acc["T7"] = int(t[0].item())
            if by_states:
                acc["T"], acc["C"] = int(t[1].item()), int(t[2].item())
            t = torch.tensor([acc["ms_filter"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            acc["ms_filter"] = float(t.item())
        acc["T7"] += acc["T7_rep"]
        acc["T"] += acc["T7"]
        acc["launches"] = eng.launches - launches0
        return ms, acc, clocks

    ms_res, acc_res, clocks = timed(resident=True)
    ms_e2e, acc_e2e, _ = timed(resident=False)

    if rank == 0:
        units = acc_res["T"] + acc_res["C"]
        value = units / (ms_res * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel: the 7-LUT phase-1 sweep, one launch per state per step
        # (per GPU: this rank's launches and the combinations they swept)
        filt_s = acc_res["ms_filter_local"] * 1e-3
        achieved = acc_res["T7_local"] * BYTES_T7 / max(filt_s, 1e-12) / 1e9
        dram_per_launch = None
        try:
            dram_per_launch = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))[
                "filter7_bytes_per_launch"]
        except (OSError, KeyError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "u32 bitwise (LOP3)",
            "data": "synthetic", "config": config,
            "t_units_per_s": acc_res["T"] / (ms_res * 1e-3),
            "c_units_per_s": acc_res["C"] / (ms_res * 1e-3),
            "units_per_step": {"T": acc_res["T"] // args.steps, "C": acc_res["C"] // args.steps},
            "kernel_ms_per_step": {k: acc_res[k] / args.steps
                                   for k in ("ms5", "ms_filter", "ms_sort", "ms_decomp")},
            "e2e": {"value": (acc_e2e["T"] + acc_e2e["C"]) / (ms_e2e * 1e-3), "unit": UNIT,
                    "ms_per_step": ms_e2e / args.steps,
                    # per state: problem block (compressed tables, target, mask; the position-major
                    # copy is derived on the device), 5-LUT position table and the two 7-LUT
                    # position tables (kernel arguments)
                    "h2d_bytes_per_step": B * (16464 + 256 + 512),
                    # per state: control words after search_5lut; 128-byte header + first 1,024 list
                    # entries after search_7lut
                    "d2h_bytes_per_step": B * (72 + 128 + 8192),
                    "note": "host tables -> sbg_load_problem -> sbg_search5/7 -> result structs"},
            "gpu_launches": acc_res["launches"],
            "roofline": {
                "bound": "hbm", "kernel": "k_filter7_pm<NW,W,P> (search_7lut phase 1)",
                "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "traffic": dram_per_launch,
                "note": "algorithmic bytes = 224 B per 7-combination (SURVEY.md 8d); operands are "
                        "served from shared memory, compulsory DRAM traffic is the problem block (31-50 KB "
                        "per launch, ncu), so frac > 1 is expected; the binding resource is integer-ALU "
                        "issue (57-71 % of that pipe's measured peak, ncu capture F), see profiles/ and "
                        "DESIGN.md",
            },
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_arm(n, B, 2000, budget_s=12.0)
            except Exception as exc:  # keep the GPU line even if the CPU leg cannot run
                line["cpu_baseline"] = {"error": repr(exc)}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
