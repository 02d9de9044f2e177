#!/usr/bin/env python
"""Times the batched (tcgen05) prefill of BASELINE.json configs[3]: TinyLlama-1.1B Q8, 2048-token synthetic prompt.

  python tools/bench_prefill.py [--workload q8|q4] [--tokens 2048] [--iters 5] [--exact-rows 0]

Prints one JSON line: prefill tokens/s (CUDA events on the library stream, H2D of the token ids and the lm_head of
the last row included), the tensor-core roofline fraction against MEASURED_PEAKS.json, and optionally the exact
row-by-row path on a few rows for comparison.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import gtb  # noqa: E402,F401
from tinyllama_cpp_b200 import weights as W  # noqa: E402


def prefill_flops(cfg, T):
    """layer linears for T rows + lm_head of the last row + causal attention."""
    per_layer = 2 * cfg.n_embd * cfg.n_embd + 2 * cfg.kv_dim * cfg.n_embd + 3 * cfg.n_ffn * cfg.n_embd
    lin = 2.0 * cfg.n_layers * per_layer * T
    head = 2.0 * cfg.n_vocab * cfg.n_embd
    attn = cfg.n_layers * cfg.n_heads * sum(2 * 2 * cfg.d_head * (i + 1) for i in range(T))
    return lin + head + attn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", choices=["q8", "q4", "f16"], default="q8")
    ap.add_argument("--tokens", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--exact-rows", type=int, default=0)
    ap.add_argument("--layers", type=int, default=0, help="debug: smaller model")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (e.g. pf_2cta=1)")
    args = ap.parse_args()
    import torch
    from tinyllama_cpp_b200 import capi
    capi.init(0)
    wdt = {"q8": W.Q8, "q4": W.Q4, "f16": W.F16}[args.workload]
    cfg = W.TINYLLAMA if not args.layers else W.mini_config(n_layers=args.layers, n_vocab=32003)
    T = args.tokens
    eng = capi.Engine(cfg, T + 128, wdt).load(W.synth_weights(cfg, wdt, seed=1))
    for o in args.opt:
        k, v = o.split("=")
        eng.set_option(k, int(v))
    prompt = W.synth_prompt(7, T, cfg.n_vocab)
    stream = torch.cuda.ExternalStream(capi.stream_handle(), device=torch.device("cuda", 0))
    for _ in range(args.warmup):
        eng.prefill_fast(prompt)
    capi.sync()
    l0 = capi.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.iters + 1)]
    ev[0].record(stream)
    for i in range(args.iters):
        eng.prefill_fast(prompt)
        ev[i + 1].record(stream)
    capi.sync()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.iters)]
    launches = (capi.launch_count() - l0) // args.iters
    best, med = min(ms), float(np.median(ms))
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"bf16_tflops": 1590.0}
    fl = prefill_flops(cfg, T)
    out = {"metric": "prefill_tokens_per_s", "workload": f"TinyLlama-1.1B {args.workload} prefill of {T} synthetic tokens (BASELINE.json configs[3])",
           "value": T / (med * 1e-3), "ms_median": med, "ms_best": best, "ms_all": ms, "launches_per_prefill": launches,
           "flops": fl, "tflops": fl / (med * 1e-3) / 1e12, "peak_tflops": peaks["bf16_tflops"],
           "frac_of_measured_tensor_peak": fl / (med * 1e-3) / 1e12 / peaks["bf16_tflops"]}
    if args.exact_rows:
        n = args.exact_rows
        eng.prefill(prompt[:n])
        capi.sync()
        t0 = time.perf_counter()
        eng.prefill(prompt[:n])
        capi.sync()
        out["exact_path_ms_per_row"] = 1e3 * (time.perf_counter() - t0) / n
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
