"""Cycle counters of the tensor-core multi-row GEMM (k_xt_gemm, CTA (0,0) of every launch, summed over one decode step of B sequences).

    python tools/xt_trace.py [--batch 64] [--variant 0]
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gtb  # noqa: E402,F401
from tinyllama_cpp_b200 import capi, weights as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--wdt", default="q4")
    a = ap.parse_args()
    wdt = W.WDTYPE_BY_NAME[a.wdt]
    capi.init(0)
    cfg = W.TINYLLAMA
    eng = capi.Engine(cfg, 512, wdt).load(W.synth_weights(cfg, wdt, seed=1))
    eng.set_option("xr_tensor", 1)
    eng.set_option("xr_variant", a.variant)
    eng.set_option("graph", 0)
    eng.batch_create(a.batch)
    eng.prefill(W.synth_prompt(100, 64, cfg.n_vocab))
    for s in range(a.batch):
        eng.batch_adopt(s)
    eng.batch_decode(2)
    eng.set_option("xr_trace", 1)
    eng.batch_decode(1)
    capi.sync()
    d = eng.read_prof(32).astype(float)
    n = max(d[0], 1)
    st = max(d[12], 1)
    print(f"batch {a.batch} variant {a.variant}: {int(n)} GEMM launches, {int(st)} stages (CTA 0,0)")
    print(f"  epilogue warp 0 : total {d[8] / st:8.0f} cycles/stage   wait tfull {d[9] / st:8.0f}   tcgen05.ld {d[10] / st:8.0f}")
    print(f"  MMA warp        : total {d[16] / st:8.0f} cycles/stage   wait full {d[20] / st:8.0f}   wait tempty {d[21] / st:8.0f}")
    o = 24
    print(f"  producer 32     : total {d[o] / st:8.0f} cycles/stage   wait empty {d[o + 1] / st:8.0f}   wait cp.async {d[o + 2] / st:8.0f}   unpack {d[o + 3] / st:8.0f}")
    eng.close()


if __name__ == "__main__":
    main()
