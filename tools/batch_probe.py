"""Aggregate token rate of the batched order-free decode (gtb_engine_batch_*) against batch size.
usage: batch_probe.py [ctx] [steps] [q4|q8]"""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import gtb  # noqa
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
wdt = {"q4": W.Q4, "q8": W.Q8}[sys.argv[3] if len(sys.argv) > 3 else "q4"]
cfg = W.TINYLLAMA
eng = capi.Engine(cfg, ctx + steps + 8, wdt).load(W.synth_weights(cfg, wdt, seed=1))
import os
eng.set_option("graph", int(os.environ.get("GRAPH", "1")))
eng.prefill_fast(W.synth_prompt(7, ctx, cfg.n_vocab))
eng.set_option("fast_decode", 1)
eng.decode(4)
capi.sync()
t0 = time.perf_counter()
eng.decode(steps)
capi.sync()
print(f"single-sequence fast_decode: {(time.perf_counter() - t0) / steps * 1e6:8.1f} us/step")
eng.set_option("fast_decode", 0)
for B in (1, 2, 4, 8, 12, 16):
    eng.batch_create(B)
    for s in range(B):
        eng.prefill_fast(W.synth_prompt(7 + s, ctx, cfg.n_vocab))
        eng.batch_adopt(s)
    eng.batch_decode(4)
    capi.sync()
    t0 = time.perf_counter()
    eng.batch_decode(steps)
    capi.sync()
    dt = (time.perf_counter() - t0) / steps
    print(f"batch {B}: {dt * 1e6:8.1f} us/step  {B / dt:9.1f} tok/s aggregate")
