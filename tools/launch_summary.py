#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share, average."""
import collections
import csv
import sys

def main(path, split_gemm=True):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    gi = 0
    ri = 0
    for x in rows:
        name = x["Kernel Name"].split("(")[0].replace("void ", "")
        if split_gemm and "k_pf_gemm" in name:
            name += " [" + ["q|k|v", "o", "gate|up", "down"][gi % 4] + "]"
            gi += 1
        if "k_xr_gemm" in name:
            targs = name[name.index("<") + 1:name.rindex(">")].split(",") if "<" in name else []
            epi = targs[1].strip() if len(targs) > 1 else ""
            if epi.replace("(int)", "") == "1":
                name += " [" + ["o", "down"][ri % 2] + "]"
                ri += 1
            else:
                name += " [" + {"0": "q|k|v", "2": "gate|up+silu", "3": "lm_head"}.get(epi.replace("(int)", ""), epi) + "]"
        agg.setdefault(name, []).append(float(x["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':48s} {'n':>4s} {'total ms':>9s} {'share':>6s} {'avg us':>8s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:48s} {len(v):4d} {sum(v) / 1e6:9.3f} {100 * sum(v) / tot:5.1f}% {sum(v) / len(v) / 1e3:8.1f}")
    print(f"{'total':48s} {len(rows):4d} {tot / 1e6:9.3f}")

if __name__ == "__main__":
    main(sys.argv[1])
