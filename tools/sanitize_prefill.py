import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import gtb
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
cfg = W.mini_config(n_layers=1, n_vocab=300)
for wdt in (W.Q4, W.F16):
    e = capi.Engine(cfg, 200, wdt).load(W.synth_weights(cfg, wdt, seed=3))
    for T in (7, 130):
        e.prefill_fast(W.synth_prompt(2, T, cfg.n_vocab))
        print(wdt, T, float(np.abs(e.read_logits()).max()))
    e.close()
