"""One 8-row decode launch of the persistent megakernel at a long context (for `ncu -k regex:k_mega`).
The context is built with the batched prefill so the set-up takes milliseconds."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import gtb  # noqa
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 1900
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = W.TINYLLAMA
eng = capi.Engine(cfg, 2048, W.Q4).load(W.synth_weights(cfg, W.Q4, seed=1))
eng.prefill_fast(W.synth_prompt(7, ctx, cfg.n_vocab))
eng.decode(2)
eng.decode(rows)
capi.sync()
print("pos", eng.position())
