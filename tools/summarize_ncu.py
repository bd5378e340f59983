"""Turn a .ncu-rep (brought back in gpurun_out/) into the tracked summaries under
profiles/: key raw metrics, opcode mix, per-CUDA-source-line instruction share.

    python tools/summarize_ncu.py gpurun_out/prof_f64_b.ncu-rep r01_f64_v2 [json_key]
"""
import csv, io, json, re, subprocess, sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rep, name = sys.argv[1], sys.argv[2]
json_key = sys.argv[3] if len(sys.argv) > 3 else None
out = ROOT / "profiles"
out.mkdir(exist_ok=True)


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__inst_executed_op_global_red.sum"]
raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, val = raw[0], raw[1], raw[2]
lines = [f"# {name}: ncu --set full --clock-control none, one launch of {val[hdr.index('Kernel Name')]}", ""]
vals = {}
for h, u, v in zip(hdr, units, val):
    if h in KEYS or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        lines.append(f"{h} [{u}] = {v}")
        vals[h] = (v, u)

sass = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "sass"))))
h2 = sass[1]
isrc, iex, ithr = h2.index("Source"), h2.index("Instructions Executed"), h2.index("Thread Instructions Executed")
mix = defaultdict(lambda: [0, 0]); tot = 0
for r in sass[2:]:
    if len(r) <= ithr: continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc].strip())
    op = (m.group(2) if m else r[isrc][:10]).split(".")[0]
    ex = int(float(r[iex] or 0)); mix[op][0] += ex; mix[op][1] += int(float(r[ithr] or 0)); tot += ex
lines += ["", f"## opcode mix (warp-instructions executed, total {tot:.4e})"]
for k, v in sorted(mix.items(), key=lambda kv: -kv[1][0])[:28]:
    lines.append(f"{k:10s} {v[0]:.3e} {100 * v[0] / tot:6.2f}%  avg active threads {v[1] / max(v[0], 1):5.1f}")

cs = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "cuda,sass"))))
agg = defaultdict(lambda: [0, 0, ""]); tot2 = 0; cur = None; hh = None
for r in cs:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hh = r; jex = hh.index("Instructions Executed"); jth = hh.index("Thread Instructions Executed"); continue
    if hh is None or r[0] == "": continue
    try: ex = int(float(r[jex] or 0)); th = int(float(r[jth] or 0))
    except (ValueError, IndexError): continue
    a = agg[(cur, int(r[0]))]; a[0] += ex; a[1] += th; a[2] = r[1].strip()[:100]; tot2 += ex
lines += ["", "## warp-instructions by CUDA source line (inlined code attributed to the callee's line)"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    lines.append(f"{k[0]:9s}:{k[1]:4d} {100 * v[0] / max(tot2, 1):6.2f}% thr {v[1] / max(v[0], 1):5.1f} | {v[2]}")
(out / f"{name}_summary.txt").write_text("\n".join(lines) + "\n")
print("\n".join(lines[:40]))

if json_key:
    def num(k):
        v, u = vals[k]; x = float(v.replace(",", ""))
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    jf = out / "ncu_render_kernel.json"
    j = json.loads(jf.read_text()) if jf.exists() else {}
    import hashlib
    hsh = hashlib.sha256()
    for nm in ("render_kernels.cuh", "path.cuh", "real.cuh", "rng.cuh", "sinks.cuh"):     # = bench.py kernel_source_sha()
        hsh.update((ROOT / "differentiable-renderer_b200" / "csrc" / nm).read_bytes())
    fp64 = sum(mix[o][0] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
    j[json_key] = {"source_sha": hsh.hexdigest()[:16],
                   "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                   "warp_instructions": tot, "fp64_warp_instructions": fp64,
                   "issue_active_pct": float(vals["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
                   "kernel_ms_under_ncu": float(vals["gpu__time_duration.sum"][0]), "source": f"profiles/{name}_summary.txt"}
    jf.write_text(json.dumps(j, indent=1) + "\n")
