"""A few steps of the batched order-free decode, eager launches (for an ncu launch list).  usage: batch_ncu.py [B] [ctx] [steps]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import gtb  # noqa
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 512
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = W.TINYLLAMA
eng = capi.Engine(cfg, ctx + 64, W.Q4).load(W.synth_weights(cfg, W.Q4, seed=1))
eng.set_option("graph", 0)
eng.batch_create(B)
for s in range(B):
    eng.prefill_fast(W.synth_prompt(7 + s, ctx, cfg.n_vocab))
    eng.batch_adopt(s)
eng.batch_decode(steps)
capi.sync()
print("pos", eng.batch_position(0))
