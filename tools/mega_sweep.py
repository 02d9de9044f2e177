"""Sweep of a runtime option of the persistent decode kernel (default: the L2 prefetch distance `pf_ahead`).

    python tools/mega_sweep.py [--opt pf_ahead] [--values 2,3,4,5,6] [--ctx 1900] [--steps 128]
"""
import argparse
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gtb  # noqa: E402,F401
from tinyllama_cpp_b200 import capi, weights as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--opt", default="pf_ahead")
    ap.add_argument("--values", default="2,3,4,5,6")
    ap.add_argument("--ctx", type=int, default=1900)
    ap.add_argument("--steps", type=int, default=128)
    a = ap.parse_args()
    capi.init(0)
    cfg = W.TINYLLAMA
    eng = capi.Engine(cfg, 2048, W.Q4).load(W.synth_weights(cfg, W.Q4, seed=1))
    prompt = W.synth_prompt(7, a.ctx, cfg.n_vocab)
    for rep in range(2):
        for v in [int(x) for x in a.values.split(",")]:
            eng.set_option(a.opt, v)
            eng.prefill_fast(prompt)
            eng.decode(8)
            capi.sync()
            t0 = time.perf_counter()
            eng.decode(a.steps)
            capi.sync()
            print(f"{a.opt} = {v}: {(time.perf_counter() - t0) * 1e3 / a.steps:.4f} ms/token", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
