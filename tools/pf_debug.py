"""Per-activation error table of the batched prefill against the CPU checker (debug helper)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import gtb  # noqa
import oracle
from oracle import Q4, Q8
from tinyllama_cpp_b200 import capi, weights as W

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

capi.init(0)
wdt = Q8 if "q8" in sys.argv else Q4
T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cfg = W.mini_config(n_layers=3, n_vocab=300)
wl = list(W.synth_weights(cfg, wdt, seed=21))
cm = oracle.best().model(cfg, 192, wdt).load(wl)
e = capi.Engine(cfg, 192, wdt).load(wl)
prompt = W.synth_prompt(5, T, cfg.n_vocab)
want = cm.logits(prompt, 0)
e.set_option("capture_acv", 1)
e.prefill_fast(prompt)
got = e.read_logits()
print("logits rel", rel(got, want), "argmax", int(np.argmax(got)), int(np.argmax(want)))
rows = [0, 1, T // 2, T - 1]
for layer in range(cfg.n_layers):
    for name, aid in oracle.LAYER_ACVS.items():
        if name == "attn_res" and layer == cfg.n_layers - 1:
            continue
        errs = []
        for row in rows:
            g, w = e.pf_acv(layer, aid, row), cm.acv(layer, aid, row)
            nz = np.count_nonzero(g != w)
            errs.append(f"{rel(g, w):.2e}/{nz}")
        print(f"L{layer}.{name:10s}", " ".join(errs))
