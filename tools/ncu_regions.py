#!/usr/bin/env python
"""Summarise an ncu report of the megakernel: headline metrics, stall mix, and instructions/samples per code region.
usage: python tools/ncu_regions.py gpurun_out/mega_v2.ncu-rep [tokens] [layers]"""
import csv, subprocess, sys, io, bisect, re
rep = sys.argv[1]
tokens = int(sys.argv[2]) if len(sys.argv) > 2 else 4
layers = int(sys.argv[3]) if len(sys.argv) > 3 else 22
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[-1]
m = dict(zip(h, v))
def g(k):
    try: return float(m[k].replace(",", ""))
    except Exception: return float("nan")
print("duration ms", g("gpu__time_duration.sum"), " warp-instr", g("smsp__inst_executed.sum"), " issue active %", g("smsp__issue_active.avg.pct_of_peak_sustained_active"))
print("dram read GB", g("dram__bytes_read.sum"), " lts bytes", m.get("lts__t_bytes.sum"), " regs", m.get("launch__registers_per_thread"))
st = {k[len("smsp__pcsamp_warps_issue_stalled_"):]: float(x) for k, x in m.items() if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")}
tot = sum(st.values())
print("stall mix:", ", ".join(f"{k} {100*x/tot:.1f}%" for k, x in sorted(st.items(), key=lambda kv: -kv[1])[:10]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
lines = []
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if r[0] != "" and len(r) > 8:
        try: lines.append((cur, int(r[0]), r[1], float(r[6]), float(r[7])))
        except Exception: pass
tot_s = sum(l[3] for l in lines); tot_i = sum(l[4] for l in lines)
# regions of gtb_mega.cuh by function start
mega = open("tinyllama.cpp_b200/csrc/gtb_mega.cuh").read().split("\n")
marks = []
for i, l in enumerate(mega):
    mm = re.match(r"^(?:template.*\n)?__device__ .*?(\w+)\(", l) or re.match(r"^__global__.*?(\w+)\(", l)
    if mm: marks.append((i + 1, mm.group(1)))
    if l.startswith("struct ") : marks.append((i + 1, "decl"))
starts = [a for a, _ in marks]
agg = {}
for f, ln, s, smp, ins in lines:
    if f == "gtb_mega.cuh":
        k = bisect.bisect_right(starts, ln) - 1
        name = "mega:" + (marks[k][1] if k >= 0 else "head")
    else:
        name = f + (":" + ("encode" if ln < 85 else "expf" if ln < 130 else "exact_sum") if f == "gtb_dev.cuh" else "")
    a = agg.setdefault(name, [0, 0]); a[0] += smp; a[1] += ins
print(f"total: {tot_i/tokens/148/layers/16:.0f} instr/warp/layer")
for k, x in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{k:34s} samples {100*x[0]/tot_s:5.1f}%  instr {100*x[1]/tot_i:5.1f}%  ({x[1]/tokens/148/layers/16:7.0f} instr/warp/layer)")
print("--- top source lines by samples")
for l in sorted(lines, key=lambda l: -l[3])[:25]:
    print(f"{l[0]}:{l[1]:4d} {100*l[3]/tot_s:4.1f}% instr {100*l[4]/tot_i:4.1f}%  {l[2].strip()[:110]}")
