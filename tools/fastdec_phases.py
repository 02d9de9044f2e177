"""Per-phase SM-clock stamps of the persistent order-free decode kernel (k_fd_mega, option "prof").
usage: fastdec_phases.py [ctx] [q4|q8] [cta]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import gtb  # noqa
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 1900
wdt = {"q4": W.Q4, "q8": W.Q8}[sys.argv[2] if len(sys.argv) > 2 else "q4"]
cta = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg = W.TINYLLAMA
eng = capi.Engine(cfg, 2048, wdt).load(W.synth_weights(cfg, wdt, seed=1))
eng.prefill_fast(W.synth_prompt(7, ctx, cfg.n_vocab))
eng.set_option("fast_decode", 1)
eng.set_option("fd_mega", 1)
eng.set_option("prof", 1)
eng.set_option("fd_prof_cta", cta)
eng.decode(8)
capi.sync()
GEMV = ["entry", "barrier", "prologue", "rows"]          # stamps: entry, after barrier, after prologue, after rows
ATTN = ["entry", "barrier", "prep+K/V staged", "blocks"] if cta < 64 and cta % 16 < 15 else ["entry", "barrier"]   # combine lands in o:entry
n = cfg.n_layers * (4 * 4 + len(ATTN)) + 4
t = eng.read_prof(n + 1).astype(np.int64)
names, layout = [], [("qkv", GEMV), ("attn", ATTN), ("o", GEMV), ("gate|up", GEMV), ("down", GEMV)]
for li in range(cfg.n_layers):
    for ph, st in layout:
        names += [f"{ph}:{s}" for s in st]
names += [f"head:{s}" for s in GEMV]
d = np.diff(t[: len(names) + 1])
# stamp i is taken at the START of segment names[i] ... the time until the next stamp belongs to the next label
seg = {}
for i in range(len(names) - 1):
    seg.setdefault(names[i + 1], []).append(d[i])
print(f"CTA {cta}, t = {eng.position()}, SM cycles (median over {cfg.n_layers} layers); 'X:entry' = tail of the previous phase up to this phase's entry")
tot = 0.0
for k, v in seg.items():
    m = float(np.median(v))
    tot += m if not k.startswith("head") else 0.0
    print(f"  {k:22s} {m:9.0f} cycles  {m / 1.9e3:6.2f} us")
print(f"  layer total            {tot:9.0f} cycles  {tot / 1.9e3:6.2f} us")
