"""Token-level contract of the order-free decode kernels at full size: teacher-forced on the reference's own greedy sequence
(tests/golden/full_q4.npz / full_q8.npz, generated from the unmodified reference), every step's row runs through fast_decode on an
EXACT K/V cache; prints the top-1 agreement rate with the golden tokens and the reference's top-1/top-2 margin at every mismatch.

    python tools/fastdec_agreement.py [full_q4|full_q8] [n_steps]
"""
import json
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import gtb  # noqa
import make_golden as MG
from tinyllama_cpp_b200 import capi, weights as W


def agreement(name="full_q4", n_steps=None):
    gold = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    wdt, n_prompt, n_new, max_ctx = MG.FULL[name]
    n_steps = n_new if n_steps is None else min(n_steps, n_new)
    cfg = W.TINYLLAMA
    e = capi.Engine(cfg, max_ctx, wdt).load(W.synth_weights(cfg, wdt, seed=1))
    gt = gold["tokens"]
    e.prefill(gt[:n_prompt])                       # exact multi-row prefill: the cache of the golden prompt
    mism, rel = [], []
    for i in range(n_steps):
        n = n_prompt + i                           # row n - 1 predicts golden token gt[n]
        e.set_option("fast_decode", 1)
        lf = e.logits(gt[:n], n - 1)
        e.set_option("fast_decode", 0)
        lx = e.logits(gt[:n], n - 1)               # the exact row: also restores the exact K/V of position n - 1
        assert int(np.argmax(lx)) == int(gt[n]), f"exact path left the golden sequence at step {i}"
        rel.append(float(np.linalg.norm(lf.astype(np.float64) - lx) / np.linalg.norm(lx.astype(np.float64))))
        if int(np.argmax(lf)) != int(gt[n]):
            srt = np.sort(lx)
            mism.append({"step": i, "ref_margin": float(srt[-1] - srt[-2]), "golden_margin": float(gold["margin"][i]),
                         "fast_rank_of_golden": int((lf > lf[gt[n]]).sum())})
    e.close()
    return {"workload": name, "steps": n_steps, "agree": n_steps - len(mism), "agreement_rate": 1 - len(mism) / n_steps,
            "max_mismatch_margin": max([m["ref_margin"] for m in mism], default=0.0), "mean_rel_l2": float(np.mean(rel)),
            "max_rel_l2": float(np.max(rel)), "mismatches": mism,
            "margin_quantiles_all_steps": [float(np.quantile(gold["margin"][:n_steps], q)) for q in (0.05, 0.25, 0.5, 0.75)]}


if __name__ == "__main__":
    capi.init(0)
    name = sys.argv[1] if len(sys.argv) > 1 else "full_q4"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else None
    print(json.dumps(agreement(name, n)))
