"""compute-sanitizer driver for the order-free decode kernels (PDL chain, persistent kernel, batched decode), 1-layer model.
usage: compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_fastdec.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import gtb  # noqa
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
cfg = W.mini_config(n_layers=1, n_vocab=300)
for wdt in (W.Q4, W.Q8):
    e = capi.Engine(cfg, 700, wdt).load(W.synth_weights(cfg, wdt, seed=3))
    e.set_option("graph", 0)
    for mega in (0, 1):
        for T in (7, 300, 600):          # 1, 2 and 3 position chunks per head
            p = W.synth_prompt(2, T, cfg.n_vocab)
            e.prefill_fast(p[:-1])
            e.set_option("fast_decode", 1)
            e.set_option("fd_mega", mega)
            e.decode(2)
            print(wdt, "mega" if mega else "pdl", T, float(np.abs(e.read_logits()).max()))
            e.set_option("fast_decode", 0)
    e.batch_create(3)
    for s, T in enumerate((5, 280, 520)):
        e.prefill_fast(W.synth_prompt(3 + s, T, cfg.n_vocab))
        e.batch_adopt(s)
    e.batch_decode(2)
    print(wdt, "batch", [e.batch_position(s) for s in range(3)], float(np.abs(e.batch_read_logits(2)).max()))
    e.batch_create(10)                   # the 16-slot instantiation of the batch kernels
    for s in range(10):
        e.prefill_fast(W.synth_prompt(30 + s, 20 + 31 * s, cfg.n_vocab))
        e.batch_adopt(s)
    e.batch_decode(2)
    print(wdt, "batch10", [e.batch_position(s) for s in range(10)], float(np.abs(e.batch_read_logits(9)).max()))
    e.close()
