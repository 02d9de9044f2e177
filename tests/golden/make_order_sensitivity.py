"""How much the REFERENCE ITSELF moves when only its summation order changes.

The reference documents two builds (README.md:18 and :25): `g++ -O3 -fopenmp` (scalar dot products, ops.h:296-312,
450-479) and `g++ -O3 -fopenmp -mavx -mf16c` (4/8-lane AVX dot products).  Both are "the reference"; they associate
the same sums differently, and every op re-encodes its output as Q8 blocks, so a last-bit difference becomes a
one-code-step difference a few ops later.  This script runs both builds of the UNMODIFIED sources
(oracle/_ref/libgten_ref.so, oracle/_ref/libgten_ref_scalar.so) on the same seeded inputs and stores their mutual
distance.  tests/test_prefill_gpu.py uses it as the yardstick for the batched (tcgen05) prefill, whose only licence
to differ from the AVX build is exactly this: summation order (plus fp16 operand rounding).

    python tests/golden/make_order_sensitivity.py mini          # seconds
    python tests/golden/make_order_sensitivity.py full          # ~15 min (two 2048-token CPU prefills of the 1.1 B model)
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import gtb  # noqa: E402,F401
import oracle  # noqa: E402
from oracle import F16, Q4, Q8  # noqa: E402
from tinyllama_cpp_b200 import weights as W  # noqa: E402

OUT = Path(__file__).resolve().parent
MINI = dict(n_layers=3, n_vocab=300, seed=21, prompt_seed=5, max_ctx=192, n_prompts=(100, 64, 7))


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def scalar_lib():
    so = oracle.HERE / "_ref" / "libgten_ref_scalar.so"
    assert so.exists(), "build it: make -C oracle ref_scalar"
    return oracle.CpuLib(so, "ref_", "reference-scalar")


def mini(out):
    avx, sca = oracle.ref(), scalar_lib()
    cfg = W.mini_config(n_layers=MINI["n_layers"], n_vocab=MINI["n_vocab"])
    for wn, wdt in (("q8", Q8), ("q4", Q4), ("f16", F16)):
        wl = list(W.synth_weights(cfg, wdt, seed=MINI["seed"]))
        for T in MINI["n_prompts"]:
            if f"mini_{wn}_{T}_logits" in out:
                continue
            a = avx.model(cfg, MINI["max_ctx"] if wdt != F16 else 2 * MINI["max_ctx"], wdt).load(wl)     # FP16 P rows: 2 n <= max_ctx
            s = sca.model(cfg, MINI["max_ctx"] if wdt != F16 else 2 * MINI["max_ctx"], wdt).load(wl)
            prompt = W.synth_prompt(MINI["prompt_seed"], T, cfg.n_vocab)
            la, ls = a.logits(prompt, 0), s.logits(prompt, 0)
            rows = sorted({0, 1, T // 2, T - 2, T - 1} & set(range(T)))
            names = list(oracle.LAYER_ACVS)
            tab = np.zeros((cfg.n_layers, len(names)))
            for layer in range(cfg.n_layers):
                for j, name in enumerate(names):
                    tab[layer, j] = max(rel(s.acv(layer, oracle.LAYER_ACVS[name], r), a.acv(layer, oracle.LAYER_ACVS[name], r)) for r in rows)
            out[f"mini_{wn}_{T}_acv"] = tab
            out[f"mini_{wn}_{T}_logits"] = rel(ls, la)
            print(f"mini {wn} T={T}: logits rel {rel(ls, la):.3e}; per-layer worst {tab.max(axis=1)}", flush=True)
            a.close(); s.close()
    out["acv_names"] = np.array(list(oracle.LAYER_ACVS))


def full(out):
    sca = scalar_lib()
    gold = np.load(OUT / "prefill_q8.npz")
    cfg = W.TINYLLAMA
    m = sca.model(cfg, 2176, Q8).load(W.synth_weights(cfg, Q8, seed=1))
    ls = m.logits(gold["tokens"][:2048], 0)
    out["full_q8_2048_logits"] = rel(ls, gold["logits"])
    out["full_q8_2048_top1_same"] = int(np.argmax(ls) == np.argmax(gold["logits"]))
    out["full_q8_2048_scalar_logits"] = ls.astype(np.float32)
    print(f"full q8 2048: scalar-vs-AVX logits rel {rel(ls, gold['logits']):.3e}; top-1 same: {np.argmax(ls) == np.argmax(gold['logits'])}", flush=True)


if __name__ == "__main__":
    f = OUT / "order_sensitivity.npz"
    out = dict(np.load(f)) if f.exists() else {}
    for what in (sys.argv[1:] or ["mini"]):
        {"mini": mini, "full": full}[what](out)
    np.savez_compressed(f, **out)
