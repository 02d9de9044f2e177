"""A few rows of the order-free decode path at a long context, eager launches (for an ncu launch list).
usage: fastdec_ncu.py [ctx] [rows] [q4|q8]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import gtb  # noqa
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 1900
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wdt = {"q4": W.Q4, "q8": W.Q8}[sys.argv[3] if len(sys.argv) > 3 else "q4"]
cfg = W.TINYLLAMA
eng = capi.Engine(cfg, 2048, wdt).load(W.synth_weights(cfg, wdt, seed=1))
eng.prefill_fast(W.synth_prompt(7, ctx, cfg.n_vocab))
eng.set_option("fast_decode", 1)
eng.set_option("graph", 0)
eng.set_option("fd_mega", 0)
eng.decode(rows)
capi.sync()
print("pos", eng.position())
