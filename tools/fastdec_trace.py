"""%globaltimer timeline of one step of the PDL chains (option "fd_trace"): single-sequence fast decode and the batched decode.
Needs the library built with the stamps: python tinyllama.cpp_b200/build.py --trace   (rebuild without it afterwards).
usage: fastdec_trace.py [ctx] [B]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import gtb  # noqa
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = W.TINYLLAMA
eng = capi.Engine(cfg, ctx + 64, W.Q4).load(W.synth_weights(cfg, W.Q4, seed=1))
KIND = {0: "norm", 1: "gemv", 2: "silu", 3: "head", 4: "attn", 7: "(wait)"}
STAGE = {0: "entry", 1: "released", 2: "end", 3: "staged"}


def dump(title, n_show=24, skip=60):
    t = eng.read_prof(4001)
    n = int(t[0] & 0xffffffff)
    ev = [(int(v) & 0xffffffffffff, int(v) >> 48) for v in t[1:1 + min(n, 4000)]]
    ev.sort()
    print(f"--- {title}: {n} events; showing {n_show} from #{skip} (us relative to the first shown)")
    t0 = ev[skip][0]
    for ns, tag in ev[skip:skip + n_show]:
        print(f"  {(ns - t0) / 1e3:8.2f}  {KIND.get(tag >> 2, tag >> 2):7s} {STAGE.get(tag & 3)}")


eng.prefill_fast(W.synth_prompt(7, ctx, cfg.n_vocab))
eng.set_option("fast_decode", 1)
eng.decode(3)
eng.set_option("fd_trace", 1)
eng.decode(1)
capi.sync()
dump("single-sequence chain")
eng.set_option("fd_trace", 0)
eng.set_option("fast_decode", 0)
eng.batch_create(B)
for s in range(B):
    eng.prefill_fast(W.synth_prompt(7 + s, ctx, cfg.n_vocab))
    eng.batch_adopt(s)
eng.batch_decode(3)
eng.set_option("fd_trace", 1)
eng.batch_decode(1)
capi.sync()
dump(f"batched chain, B = {B}", n_show=32, skip=63)
