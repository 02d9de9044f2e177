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

LABELS = ["P1 gather+norm (+embed)", "P1 gemv q|k|v", "P2a rope+scores", "P2b softmax+P.V", "P3 gather+encode", "P3 gemv o",
          "P4 gather+norm", "P4 gemv gate|up", "P4b silu*up", "P5 gather act", "P5 gemv down"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="q4")
    ap.add_argument("--ctx", type=int, default=1900)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--pf", type=int, default=-1)
    ap.add_argument("--layers", type=int, default=22)
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
    L = cfg.n_layers
    n = 1 + 11 * L + 2
    acc = np.zeros(11)
    head = np.zeros(2)
    tot = 0.0
    for _ in range(a.steps):
        eng.decode(1)
        t = eng.read_prof(n).astype(np.float64)
        d = np.diff(t)
        acc += d[: 11 * L].reshape(L, 11).mean(axis=0)
        head += d[11 * L: 11 * L + 2]
        tot += t[n - 1] - t[0]
    acc /= a.steps
    head /= a.steps
    tot /= a.steps
    print(f"workload {a.workload} ctx {a.ctx}: row {tot / 1e3:.1f} us (with prof stamps)")
    for i, nm in enumerate(LABELS):
        print(f"  {nm:26s} {acc[i]:9.0f} ns/layer")
    print(f"  per-layer sum              {acc.sum():9.0f} ns  x {L} = {acc.sum() * L / 1e3:.1f} us")
    print(f"  head prologue {head[0]:.0f} ns, lm_head+argmax {head[1]:.0f} ns")


if __name__ == "__main__":
    main()
