"""Static SASS instructions per CUDA source line for one kernel (development aid).
nvdisasm -g attributes inlined code to the innermost line; for the straight-line
segment body of the render kernels the static count per line is the executed
count per segment.  usage: sass_lines.py <lib.so> <substring of mangled name> [file-substring]"""
import re, subprocess, sys, tempfile, os, glob
from collections import Counter, defaultdict
lib, pat = sys.argv[1], sys.argv[2]
filt = sys.argv[3] if len(sys.argv) > 3 else ""
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = glob.glob(tmp + "/*.cubin")[0]
out = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
infun, cur = False, None
per = defaultdict(Counter)
for line in out.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m: infun = pat in m.group(1); continue
    if not infun: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur: per[cur][m.group(2)] += 1
F64 = ("DFMA", "DMUL", "DADD", "DSETP")
tot = slots = 0
rows = []
for k, c in per.items():
    n = sum(c.values()); f = sum(c[o] for o in F64)
    tot += n; slots += n + f
    if filt in k[0]: rows.append((k, n, f, c))
print(f"total {tot} slots {slots}")
for k, n, f, c in sorted(rows):
    print(f"{k[0]}:{k[1]:4d}  n {n:4d} fp64 {f:3d}  " + " ".join(f"{o}{v}" for o, v in c.most_common(6)))
