"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libgten_ref.so, compiled from
/root/reference with its own `-O3 -fopenmp -mavx -mf16c` line).  Run in the build container only:

    python tests/golden/make_golden.py ops mini            # seconds
    python tests/golden/make_golden.py full_f16 full_q8    # ~2 min each (1.1 B synthetic weights + reference run)
    python tests/golden/make_golden.py full_q4 prefill_q8  # ~5 min each

Inputs are regenerated from seeds by the tests (tinyllama_cpp_b200.weights is platform independent), so only
outputs are stored.  The reference publishes no golden vectors of its own.
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import gtb  # noqa: E402,F401
import oracle  # noqa: E402
from oracle import F16, F32, Q4, Q8  # noqa: E402
from tinyllama_cpp_b200 import weights as W  # noqa: E402

OUT = Path(__file__).resolve().parent
ADT = {F16: F16, Q8: Q8, Q4: Q8}

# (name, wdtype, n_prompt, n_new, max_ctx): BASELINE.json configs 1-3
FULL = {
    "full_f16": (F16, 128, 256, 384),
    "full_q8": (Q8, 128, 256, 384),
    "full_q4": (Q4, 1536, 512, 2048),
}


SEQ64 = dict(n_seq=4, n_prompt=128, n_new=64, max_ctx=192)


def op_inputs(seed=1234):
    """Seeded op-level inputs shared by make_golden and the tests."""
    rng = np.random.default_rng(seed)
    d = {}
    d["x3"] = (rng.standard_normal((3, 2048)) * 1.7).astype(np.float32)
    d["y3"] = (rng.standard_normal((3, 2048)) * 0.6).astype(np.float32)
    d["xf"] = (rng.standard_normal((2, 5632)) * 1.1).astype(np.float32)
    d["w_small"] = (rng.standard_normal((64, 2048)) * 0.02).astype(np.float32)
    d["w_down"] = (rng.standard_normal((32, 5632)) * 0.02).astype(np.float32)
    d["normw"] = (1 + 0.1 * rng.standard_normal(2048)).astype(np.float16)
    d["q"] = rng.standard_normal((9, 2048)).astype(np.float32)
    d["k"] = rng.standard_normal((9, 256)).astype(np.float32)
    d["v"] = rng.standard_normal((9, 256)).astype(np.float32)
    d["tokens"] = rng.integers(0, 40, 5).astype(np.int32)
    d["emb"] = (rng.standard_normal((40, 2048)) * 0.02).astype(np.float32)
    return d


def run_ops(lib):
    """Every op of §8(a) on the seeded inputs, through one CPU checker; returns {name: uint8/float array}."""
    I = op_inputs()
    out = {}
    for adt, an in ((Q8, "q8"), (F16, "f16")):
        x, y = lib.encode_rows(I["x3"], adt), lib.encode_rows(I["y3"], adt)
        out[f"enc_{an}"] = x
        out[f"dec_{an}"] = lib.decode_rows(x, adt, 2048)
        out[f"norm_{an}"] = lib.rms_norm(x, adt, 3, 2048, I["normw"])
        out[f"silu_{an}"] = lib.silu(x, adt, 3, 2048)
        out[f"mul_{an}"] = lib.mul(x, y, adt, 3, 2048)
        out[f"add_{an}"] = lib.add(x, y, adt, 3, 2048)
        out[f"rope_{an}"] = lib.rotary_emb(x, adt, 3, 2048, 64)
        q, k, v = lib.encode_rows(I["q"], adt), lib.encode_rows(I["k"], adt), lib.encode_rows(I["v"], adt)
        out[f"attn_{an}"] = lib.qkv_attn(q, k, v, adt, 9, 32, 4, 64, 64)
        out[f"attn_last_{an}"] = lib.qkv_attn(q, k, v, adt, 9, 32, 4, 64, 64, start_pos=8)[8:]
    for wdt, wn in ((Q4, "q4"), (Q8, "q8"), (F16, "f16")):
        adt = ADT[wdt]
        x = lib.encode_rows(I["x3"], adt)
        xf = lib.encode_rows(I["xf"], adt)
        w = W.quantize_payload(I["w_small"], wdt)
        wd = W.quantize_payload(I["w_down"], wdt)
        out[f"matmul_{wn}"] = lib.matmul_2d(x, adt, 3, 2048, w, wdt, 64, adt)
        out[f"matmul_down_{wn}"] = lib.matmul_2d(xf, adt, 2, 5632, wd, wdt, 32, adt)
        out[f"logits_{wn}"] = lib.matmul_2d(x, adt, 3, 2048, w, wdt, 64, F32, out_1d=True, start_pos=2)
        emb = W.quantize_payload(I["emb"], wdt)
        out[f"embed_{wn}"] = lib.token_embed(emb, wdt, 40, 2048, I["tokens"], adt)
        out[f"wdeq_{wn}"] = np.stack([lib.read_row(r, wdt, 2048) for r in w.reshape(64, -1)])
    return out


MINI = dict(n_layers=2, n_vocab=256, seed=3, prompt_seed=11, n_prompt=20, n_new=12, max_ctx=96)


def run_mini(lib, wdt):
    cfg = W.mini_config(n_layers=MINI["n_layers"], n_vocab=MINI["n_vocab"])
    m = lib.model(cfg, MINI["max_ctx"], wdt).load(W.synth_weights(cfg, wdt, seed=MINI["seed"]))
    prompt = W.synth_prompt(MINI["prompt_seed"], MINI["n_prompt"], cfg.n_vocab)
    toks, _, lg = m.generate(prompt, MINI["n_new"], want_logits=True)
    row = MINI["n_prompt"] + MINI["n_new"] - 2
    acv = {f"L{l}_{n}": m.acv(l, a, row) for l in range(cfg.n_layers) for n, a in oracle.LAYER_ACVS.items()}
    m.close()
    return toks, lg, acv


def summarize_logits(lg):
    """Per-step top-1 value, top-1/top-2 margin and a float64 checksum."""
    srt = np.sort(lg, axis=1)
    return srt[:, -1].copy(), (srt[:, -1] - srt[:, -2]).copy(), lg.astype(np.float64).sum(axis=1)


def run_full(lib, wdt, n_prompt, n_new, max_ctx, seed=1, prompt_seed=7):
    cfg = W.TINYLLAMA
    t0 = time.time()
    m = lib.model(cfg, max_ctx, wdt).load(W.synth_weights(cfg, wdt, seed=seed))
    print(f"  weights ready in {time.time() - t0:.0f}s", flush=True)
    prompt = W.synth_prompt(prompt_seed, n_prompt, cfg.n_vocab)
    toks, times, lg = m.generate(prompt, n_new, want_logits=True)
    m.close()
    return toks, times, lg


def main(argv):
    lib = oracle.ref()
    print(lib.build_info())
    for what in argv:
        t0 = time.time()
        if what == "ops":
            np.savez_compressed(OUT / "ops.npz", **run_ops(lib))
        elif what == "mini":
            d = {}
            for wdt, wn in ((Q4, "q4"), (Q8, "q8"), (F16, "f16")):
                toks, lg, acv = run_mini(lib, wdt)
                d[f"{wn}_tokens"] = toks
                d[f"{wn}_logits"] = lg
                for k, v in acv.items():
                    d[f"{wn}_{k}"] = v
            np.savez_compressed(OUT / "mini.npz", **d)
        elif what in FULL:
            wdt, n_prompt, n_new, max_ctx = FULL[what]
            toks, times, lg = run_full(lib, wdt, n_prompt, n_new, max_ctx)
            top1, margin, csum = summarize_logits(lg)
            keep = [0, 1, n_new // 2, n_new - 1]
            np.savez_compressed(OUT / f"{what}.npz", tokens=toks, top1=top1, margin=margin, checksum=csum,
                                keep_steps=np.array(keep), keep_logits=lg[keep],
                                cpu_prefill_s=times[0], cpu_decode_s=times[1], n_prompt=n_prompt, n_new=n_new, max_ctx=max_ctx)
            print(f"  reference CPU here: prefill {n_prompt / times[0]:.1f} tok/s, decode {(n_new - 1) / times[1]:.2f} tok/s")
        elif what == "seq64_q4":
            # BASELINE.json config 5 (bench.py --workload q4_seq64): the first SEQ64["n_seq"] of the 64 sequences, 128 + 64 tokens each
            S = SEQ64
            cfg = W.TINYLLAMA
            m = lib.model(cfg, S["max_ctx"], Q4).load(W.synth_weights(cfg, Q4, seed=1))
            toks, last, csum = [], [], []
            for sidx in range(S["n_seq"]):
                t, _, lg = m.generate(W.synth_prompt(100 + sidx, S["n_prompt"], cfg.n_vocab), S["n_new"], want_logits=True)
                toks.append(t); last.append(lg[-1].copy()); csum.append(lg.astype(np.float64).sum(axis=1))
                print(f"  sequence {sidx} done", flush=True)
            m.close()
            np.savez_compressed(OUT / "seq64_q4.npz", tokens=np.stack(toks), last_logits=np.stack(last), checksum=np.stack(csum))
        elif what == "prefill_q8":
            # BASELINE.json config 4: Q8 prefill of 2048 tokens; the oracle needs max_ctx >= 2176
            toks, times, lg = run_full(lib, Q8, 2048, 1, 2176)
            np.savez_compressed(OUT / "prefill_q8.npz", tokens=toks, logits=lg[0], cpu_prefill_s=times[0], n_prompt=2048, max_ctx=2176)
            print(f"  reference CPU here: prefill {2048 / times[0]:.1f} tok/s")
        else:
            raise SystemExit(f"unknown target {what}")
        print(f"{what}: done in {time.time() - t0:.0f}s", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["ops", "mini"])
