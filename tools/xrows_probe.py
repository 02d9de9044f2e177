"""Timing probe of the order-exact multi-row path (gtb_xrows.cu) at full size: exact prefill and exact batched decode.

    python tools/xrows_probe.py [--wdt q4|q8] [--prompt 1536] [--seqs 64] [--seq-prompt 128] [--steps 32] [--batches 64,32,16,8]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gtb  # noqa: E402,F401
from tinyllama_cpp_b200 import capi, weights as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--wdt", default="q4")
    ap.add_argument("--prompt", type=int, default=1536)
    ap.add_argument("--seq-prompt", type=int, default=128)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--batches", default="64,32,16,8")
    ap.add_argument("--max-ctx", type=int, default=2048)
    ap.add_argument("--skip-prefill", action="store_true")
    ap.add_argument("--adopt", action="store_true", help="one prefill on the engine's own sequence, copied into every slot (few launches: for ncu)")
    ap.add_argument("--prefill-reps", type=int, default=2)
    ap.add_argument("--pdl", type=int, default=1)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--tensor", type=int, default=0, help="Linears on the tensor cores (1) or the SIMT dp4a kernel (0)")
    args = ap.parse_args()
    wdt = W.WDTYPE_BY_NAME[args.wdt if args.wdt != "f16" else "fp16"]
    capi.init(0)
    cfg = W.TINYLLAMA
    eng = capi.Engine(cfg, args.max_ctx, wdt).load(W.synth_weights(cfg, wdt, seed=1))
    out = {}
    eng.set_option("xr_pdl", args.pdl)
    eng.set_option("xr_variant", args.variant)
    eng.set_option("xr_tensor", args.tensor)
    if not args.skip_prefill:
        prompt = W.synth_prompt(7, args.prompt, cfg.n_vocab)
        if args.prefill_reps > 1:
            eng.prefill(prompt[:70])
        capi.sync()
        for rep in range(args.prefill_reps):
            t0 = time.perf_counter()
            eng.prefill(prompt)
            capi.sync()
            ms = (time.perf_counter() - t0) * 1e3
        out["exact_prefill"] = {"tokens": args.prompt, "ms": ms, "tok_s": args.prompt / ms * 1e3, "first_token": int(eng.read_tokens(args.prompt, 1)[0])}
        if args.prefill_reps > 1:
            eng.set_option("xrows", 0)
            t0 = time.perf_counter()
            eng.prefill(prompt[:256])
            capi.sync()
            ms1 = (time.perf_counter() - t0) * 1e3
            eng.set_option("xrows", 1)
            out["row_at_a_time_prefill_256"] = {"ms": ms1, "tok_s": 256 / ms1 * 1e3}
        print(json.dumps(out), flush=True)
    for B in [int(x) for x in args.batches.split(",") if x]:
        eng.batch_create(B)
        t0 = time.perf_counter()
        if args.adopt:
            eng.prefill(W.synth_prompt(100, args.seq_prompt, cfg.n_vocab))
        for s in range(B):
            if args.adopt:
                eng.batch_adopt(s)
            else:
                eng.batch_prefill(s, W.synth_prompt(100 + s, args.seq_prompt, cfg.n_vocab))
        capi.sync()
        pf_ms = (time.perf_counter() - t0) * 1e3
        eng.batch_decode(2)                     # graph capture + warm-up
        capi.sync()
        l0 = capi.launch_count()
        t0 = time.perf_counter()
        eng.batch_decode(args.steps)
        capi.sync()
        ms = (time.perf_counter() - t0) * 1e3
        out[f"batch{B}"] = {"prefill_ms": pf_ms, "prefill_tok_s": B * args.seq_prompt / pf_ms * 1e3, "ms_per_step": ms / args.steps,
                            "tok_s": B * args.steps / ms * 1e3, "launches_per_step": (capi.launch_count() - l0) / args.steps,
                            "t_range": [args.seq_prompt + 3, args.seq_prompt + 2 + args.steps]}
        print(json.dumps({f"batch{B}": out[f"batch{B}"]}), flush=True)
    eng.batch_create(0)
    eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
