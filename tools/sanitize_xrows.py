"""compute-sanitizer driver for the order-exact multi-row kernels (gtb_xrows.cu) and the round-2 k_mega changes, 1-layer models.
usage: compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_xrows.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import gtb  # noqa
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
cfg = W.mini_config(n_layers=1, n_vocab=300)
QUICK = "--quick" in sys.argv               # racecheck is slow: one quantised and the fp16 model, fewer prompt lengths
TENSOR = "--tensor" in sys.argv             # the xr_tensor experiment (tcgen05 Linears) instead of the SIMT GEMM
for wdt in ((W.Q4, W.F16) if QUICK else (W.Q4, W.Q8, W.F16)):
    e = capi.Engine(cfg, 400, wdt).load(W.synth_weights(cfg, wdt, seed=3))
    e.set_option("graph", 0)
    e.set_option("xr_tensor", 1 if (TENSOR and wdt != W.F16) else 0)
    for T in ((5, 70) if QUICK else (5, 40, 70, 150)):                  # per-head attention (<= 32 rows per pass), group attention (64 rows), tails
        toks = e.generate(W.synth_prompt(2, T, cfg.n_vocab), 3)      # multi-row prefill, then k_mega (K/V copies transposed first)
        print(wdt, "prefill", T, toks[-3:].tolist(), flush=True)
    e.batch_create(40)                          # > 32 rows: the group attention kernel in decode mode
    for s in range(40):
        e.batch_prefill(s, W.synth_prompt(30 + s, 3 + 4 * s, cfg.n_vocab))
    e.batch_decode(2)
    print(wdt, "batch40", [e.batch_position(s) for s in (0, 39)], float(np.abs(e.batch_read_logits(39)).max()), flush=True)
    e.batch_create(5)                           # <= 16 rows: programmatic dependent launch + per-head attention
    for s in range(5):
        e.batch_prefill(s, W.synth_prompt(3 + s, 20 + 60 * s, cfg.n_vocab))
    e.batch_decode(2)
    print(wdt, "batch5", [e.batch_position(s) for s in range(5)], flush=True)
    e.close()
