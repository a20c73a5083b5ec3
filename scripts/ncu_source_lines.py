"""Per-source-line instruction and stall-sample shares from `ncu -i X.ncu-rep --page source --csv
--print-source cuda,sass > file.csv`.  usage: ncu_source_lines.py file.csv [kernel-substring] [top]"""
import csv
import sys

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
csv.field_size_limit(1 << 30)
out, fn, fpath, hdr = {}, None, None, None
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        fpath, hdr = r[1], None
        continue
    if r[0] == "Function Name":
        fn, hdr = r[1], None
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0] != "" and fpath.endswith("sbg_device.cuh") and want in fn:
        L, ie, si = len(hdr), hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            key = (fn, int(r[0]))
            src = ",".join(r[1:len(r) - (L - 2)])
            out[key] = (src, int(r[len(r) - (L - ie)]), int(r[len(r) - (L - si)]))
        except ValueError:
            pass
fns = sorted({k[0] for k in out})
for f in fns:
    rows = [(k[1],) + v for k, v in out.items() if k[0] == f]
    ti, ts = sum(x[2] for x in rows), sum(x[3] for x in rows)
    print("==", f[:90], "instr", ti, "samples", ts)
    for ln, src, c, sm in sorted(rows, key=lambda x: -x[3])[:top]:
        print("%5d %9d %5.1f%% smp %5d %4.1f%% | %s" % (ln, c, 100.0 * c / max(ti, 1), sm,
                                                       100.0 * sm / max(ts, 1), src[:100]))
