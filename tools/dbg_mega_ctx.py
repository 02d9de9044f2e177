"""Debug: row-at-a-time persistent kernel vs the multi-row path on long contexts of a mini model."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gtb  # noqa
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
cfg = W.mini_config(n_layers=2, n_vocab=300)
wl = list(W.synth_weights(cfg, W.Q4, seed=8))
for max_ctx, n in ((720, 700), (768, 700), (1024, 700), (2048, 700), (720, 500), (720, 300), (600, 520), (720, 640), (720, 520)):
    p = W.synth_prompt(70, n, cfg.n_vocab)
    a = capi.Engine(cfg, max_ctx, W.Q4).load(wl)
    ta = a.generate(p, 10)
    b = capi.Engine(cfg, max_ctx, W.Q4).load(wl)
    b.set_option("xrows", 0)
    tb = b.generate(p, 10)
    b.prefill(p); b.decode(9)
    tc = b.read_tokens(0, n + 10)
    b.set_option("mega", 0)
    td = b.generate(p, 10)
    print(max_ctx, n, "xr:", ta[n:].tolist(), "mega-generate:", tb[n:].tolist(), "mega prefill+decode:", tc[n:].tolist(), "phase kernels:", td[n:].tolist(), flush=True)
    a.close(); b.close()
