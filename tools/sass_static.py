"""Static SASS opcode count of one kernel of libdrtb.so (development aid): the
render kernels are issue bound, so the instruction count of the straight-line
segment body tracks their run time.  usage: sass_static.py <lib.so> <substring of the mangled name>"""
import re, subprocess, sys
from collections import Counter
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, ops = None, Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m: ops[m.group(2)] += 1
tot = sum(ops.values())
fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"total {tot}  fp64 {fp64}  slots(fp64 x2) {tot + fp64}")
print("  ".join(f"{k} {v}" for k, v in ops.most_common(24)))
