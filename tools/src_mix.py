"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA
source line (development aid).  Files are printed one after another; each has
'File Path' then a header row, then rows: line rows (Line No, Source, '', ...) and
sass rows."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
cur_file = None; hdr = None; cur_line = None; cur_src = ""
agg = defaultdict(lambda: [0, 0, 0, ""])
tot = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; iex = hdr.index("Instructions Executed"); ithr = hdr.index("Thread Instructions Executed"); ismp = hdr.index("# Samples"); continue
    if hdr is None: continue
    if r[0] != "":                      # a CUDA source line row (carries aggregated metrics for the line)
        try:
            ex = int(float(r[iex] or 0)); th = int(float(r[ithr] or 0)); sm = int(float(r[ismp] or 0))
        except (ValueError, IndexError):
            continue
        key = (cur_file, int(r[0]))
        agg[key][0] += ex; agg[key][1] += th; agg[key][2] += sm; agg[key][3] = r[1].strip()[:110]
        tot += ex
print(f"total {tot:.3e}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 45
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{k[0]:10s}:{k[1]:4d} {100*v[0]/tot:6.2f}% thr {v[1]/max(v[0],1):5.1f} smp {v[2]:7d} | {v[3]}")
