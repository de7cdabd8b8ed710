"""Aggregates an `ncu --page source --csv --print-source sass,cuda` dump per CUDA source line:
instructions executed, average active threads, stall samples.  Usage: ncu -i x.ncu-rep --page source --csv --print-source sass,cuda | python tools/ncu_lines.py [top]"""
import csv, sys, collections
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
files = {}
cur_file = None
hdr = None
agg = collections.OrderedDict()
cur_line = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    ln = r[0]
    if ln != "":
        cur_line = (cur_file, ln, r[1].strip()[:90])
        d = agg.setdefault(cur_line, [0, 0, 0, 0])
        ie = hdr.index("Instructions Executed"); te = hdr.index("Thread Instructions Executed"); sm = hdr.index("# Samples")
        try:
            d[0] += int(r[ie]); d[1] += int(r[te]); d[2] += int(r[sm])
        except ValueError:
            pass
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[2] for v in agg.values()) or 1
print("total warp-inst %d, thread-inst/warp-inst %.2f, samples %d" % (tot_i, sum(v[1] for v in agg.values()) / tot_i, tot_s))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  thr/inst %5.1f  %s:%s  %s" % (100.0 * v[0] / tot_i, 100.0 * v[2] / tot_s, (v[1] / v[0]) if v[0] else 0, k[0], k[1], k[2]))
