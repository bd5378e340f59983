"""Aggregate an `ncu --page source --print-source sass --csv` dump by opcode
(development aid): where do the executed warp-instructions go?"""
import csv, sys, re
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ithr, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
d = defaultdict(lambda: [0, 0, 0])
tot = 0
for r in rows[2:]:
    if len(r) <= ithr: continue
    src = r[isrc].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src[:12]
    key = op.split(".")[0] if "--full" not in sys.argv else op
    ex = int(float(r[iex] or 0)); th = int(float(r[ithr] or 0)); sm = int(float(r[ismp] or 0))
    d[key][0] += ex; d[key][1] += th; d[key][2] += sm; tot += ex
print(f"total warp-instructions {tot:.3e}")
for k, v in sorted(d.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{k:14s} {v[0]:.3e} {100*v[0]/tot:6.2f}%  avg-threads {v[1]/max(v[0],1):5.1f}  samples {v[2]}")
