#!/usr/bin/env python
"""Per-phase time breakdown of one decode row of the persistent megakernel (globaltimer stamps of CTA 0).

  python tools/prof_phases.py [--workload q4|q8|f16] [--ctx N] [--steps K]

Prefills N synthetic tokens, then decodes K tokens with the "prof" option on and prints, per phase of a layer,
the mean time over layers and over the K rows (ns).  Diagnostic tool: numbers taken with profiling stamps on.
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gtb  # noqa: E402,F401
from tinyllama_cpp_b200 import capi, weights as W  # noqa: E402

KIND = {0: "P1 q|k|v", 1: "P3 o", 2: "P4 gate|up", 3: "P5 down", 4: "head", 5: "  P2b", 6: "  P2a"}
GEMV = {0: "gemv: products of the tile", 1: "gemv: next tile loads issued", 2: "gemv: barrier", 3: "gemv: ordered chain"}
SUB = {80: "scores loaded (LL words)", 81: "max over the row", 82: "expf x 4", 83: "exact in-order sum", 84: "p = e / sum, re-encode, publish to smem",
       85: "P.V lane chains", 96: "V slice arrived, decoded, staged", 99: "entry barrier + K row loads issued", 100: "V loads issued", 97: "wait: this row's q, k, v", 98: "E / RoPE / E of q, k, v"}
STEP = {0: "wait for input (exchange)", 1: "LL loads + residual/encodes + squares", 2: "exact in-order sum", 3: "normalise/encode/stage (prologue done)",
        4: "gemv (products, chain, publish)", 5: "P2a: q/k/v encode, rope, scores", 6: "P2: wait scores", 7: "P2b: softmax, P.V",
        8: "P4b: silu*up", 9: "argmax exchange", 10: "L2 prefetch instructions issued (thread 0)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="q4")
    ap.add_argument("--ctx", type=int, default=1900)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--pf", type=int, default=-1)
    ap.add_argument("--layers", type=int, default=22)
    ap.add_argument("--cta", type=int, default=0, help="which CTA writes the stamps")
    a = ap.parse_args()
    wdt = W.WDTYPE_BY_NAME[{"f16": "fp16"}.get(a.workload, a.workload)]
    capi.init(0)
    cfg = W.TINYLLAMA if a.layers == 22 else W.mini_config(n_layers=a.layers, n_vocab=32003)
    eng = capi.Engine(cfg, a.ctx + a.steps + 8, wdt).load(W.synth_weights(cfg, wdt, seed=1))
    if a.pf >= 0:
        eng.set_option("pf_ahead", a.pf)
    eng.prefill(W.synth_prompt(7, a.ctx, cfg.n_vocab))
    eng.decode(4)
    eng.set_option("prof", 1)
    eng.set_option("prof_cta", a.cta)
    L = cfg.n_layers
    agg = {}
    tot = 0.0
    for _ in range(a.steps):
        eng.decode(1)
        t = eng.read_prof(4096)
        codes, times = t[0::2], t[1::2].astype(np.float64)
        n = 1
        while n < codes.size and codes[n] != 255 and times[n] >= times[n - 1] and times[n] > 0:
            n += 1
        for i in range(1, n):
            agg.setdefault(int(codes[i]), []).append(times[i] - times[i - 1])
        tot += times[n - 1] - times[0]
    tot /= a.steps
    print(f"workload {a.workload} ctx {a.ctx} cta {a.cta}: row {tot / 1e3:.1f} us (with prof stamps), {L} layers")
    total = 0.0
    for code in sorted(agg):
        v = np.array(agg[code])
        per_row = v.sum() / a.steps
        total += per_row
        if code >= 128:
            kname, sname = "  " + KIND[(code - 128) >> 3], GEMV[(code - 128) & 7]
        else:
            kname, sname = KIND[code >> 4], SUB.get(code, STEP.get(code & 15, '?'))
        print(f"  {kname:11s} {sname:46s} mean {v.mean():8.0f} ns  x{len(v) // a.steps:3d} = {per_row / 1e3:7.1f} us/row")
    print(f"  sum {total / 1e3:.1f} us")


if __name__ == "__main__":
    main()
