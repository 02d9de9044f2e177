"""Debug: replicate tests/test_xrows_gpu.py::test_exact_batch_equals_single_sequence_long_context with variations."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gtb  # noqa
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
cfg = W.mini_config(n_layers=2, n_vocab=300)
wdt = W.Q4
wl = list(W.synth_weights(cfg, wdt, seed=8))
lens = (700, 333, 256, 257, 1)
steps = 9
prompts = [W.synth_prompt(70 + i, n, cfg.n_vocab) for i, n in enumerate(lens)]
for variant in ("plain", "sync-before-one", "no-batch", "one-first"):
    one = None
    if variant == "one-first":
        one = capi.Engine(cfg, 720, wdt).load(wl)
    e = capi.Engine(cfg, 720, wdt).load(wl)
    if variant != "no-batch":
        e.batch_create(len(lens))
        for s, p in enumerate(prompts):
            e.batch_prefill(s, p)
        e.batch_decode(steps)
    if variant == "sync-before-one":
        capi.sync()
    if one is None:
        one = capi.Engine(cfg, 720, wdt).load(wl)
    one.set_option("xrows", 0)
    for s, p in enumerate(prompts):
        one.prefill(p)
        one.decode(steps)
        n = len(p)
        got = one.read_tokens(0, n + steps + 1)[n:]
        ref = e.batch_read_tokens(s, 0, n + steps + 1)[n:] if variant != "no-batch" else None
        print(variant, s, "one:", got.tolist(), "batch:", None if ref is None else ref.tolist(), flush=True)
    one.close(); e.close()
