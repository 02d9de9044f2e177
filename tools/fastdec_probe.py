"""Times the order-free decode path (option fast_decode) at a long context and compares its logits with the order-exact
path on the same K/V cache.  usage: fastdec_probe.py [ctx] [steps] [q4|q8]"""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import gtb  # noqa
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 1900
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
wdt = {"q4": W.Q4, "q8": W.Q8}[sys.argv[3] if len(sys.argv) > 3 else "q4"]
cfg = W.TINYLLAMA
prompt = W.synth_prompt(7, ctx, cfg.n_vocab)
eng = capi.Engine(cfg, 2048, wdt).load(W.synth_weights(cfg, wdt, seed=1))
eng.prefill_fast(prompt)
first = int(eng.read_tokens(ctx, 1)[0])
toks = np.concatenate([prompt, [first]]).astype(np.int32)
exact = eng.logits(toks, ctx)
eng.set_option("fast_decode", 1)
fast = eng.logits(toks, ctx)
err = float(np.linalg.norm(fast.astype(np.float64) - exact) / np.linalg.norm(exact.astype(np.float64)))
print(f"logits at t={ctx + 1}: fast vs exact rel L2 {err:.3e}; top-1 {int(np.argmax(fast))} vs {int(np.argmax(exact))}")
bytes_tok = cfg.decode_bytes(wdt, ctx + steps // 2)
cases = [("exact megakernel", {"fast_decode": 0}), ("fast, pdl eager", {"fast_decode": 1, "fd_mega": 0, "graph": 0}),
         ("fast, pdl graph", {"fast_decode": 1, "fd_mega": 0, "graph": 1})]
cases += [(f"fast, pdl graph, ahead {d}", {"fast_decode": 1, "fd_mega": 0, "graph": 1, "fd_ahead": d}) for d in (0, 6)]
cases += [(f"fast, persistent, ahead {d}", {"fast_decode": 1, "fd_mega": 1, "fd_ahead": d}) for d in (0, 3)]
for label, opts in cases:
    for k, v in opts.items():
        eng.set_option(k, v)
    eng.prefill_fast(prompt)
    eng.decode(4)
    capi.sync()
    t0 = time.perf_counter()
    eng.decode(steps)
    capi.sync()
    dt = (time.perf_counter() - t0) / steps
    print(f"{label:32s}: {dt * 1e6:8.1f} us/token  {1 / dt:8.1f} tok/s  {bytes_tok / dt / 1e9:7.1f} GB/s algorithmic")
