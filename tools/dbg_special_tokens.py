"""Debug: prompts with the chat template's special ids (32000..32002, the last rows of the embedding table)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gtb  # noqa
import oracle
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
cfg = W.mini_config(n_layers=2, n_vocab=32003)
prompt = np.array([1, 32001, 1404, 13, 22110, 338, 8425, 28579, 29973, 32002, 29871, 13, 32001, 20255, 13], np.int32)
for wdt in (W.Q8, W.Q4, W.F16):
    wl = list(W.synth_weights(cfg, wdt, seed=1))
    cm = oracle.best().model(cfg, 64, wdt).load(wl)
    want = cm.generate(prompt, 6)[0]
    e = capi.Engine(cfg, 64, wdt).load(wl)
    a = e.generate(prompt, 6)
    e.set_option("xrows", 0)
    b = e.generate(prompt, 6)
    e.set_option("mega", 0)
    c = e.generate(prompt, 6)
    print(wdt, "ref", want[15:].tolist(), "xr", a[15:].tolist(), "mega", b[15:].tolist(), "phase", c[15:].tolist(), flush=True)
    e.close(); cm.close()
