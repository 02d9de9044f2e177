"""Golden vectors (tests/golden/*.npz, produced from the UNMODIFIED reference by make_golden.py) checked
against (a) the plain-C restatement on the CPU and (b) the CUDA path through the C-ABI on a GPU."""
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle
from oracle import F16, Q4, Q8
from tinyllama_cpp_b200 import weights as W

GOLD = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(GOLD))
import make_golden as MG  # noqa: E402

WDT = {"q4": Q4, "q8": Q8, "f16": F16}


def _eq(a, b, name):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, name
    if a.dtype.kind == "f":
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), name
    else:
        assert np.array_equal(a, b), name


def test_port_matches_golden_ops(port):
    gold = np.load(GOLD / "ops.npz")
    got = MG.run_ops(port)
    assert set(got) == set(gold.files)
    for k in gold.files:
        _eq(got[k], gold[k], k)


@pytest.mark.parametrize("wn", ["q4", "q8", "f16"])
def test_port_matches_golden_mini(port, wn):
    gold = np.load(GOLD / "mini.npz")
    toks, lg, acv = MG.run_mini(port, WDT[wn])
    _eq(toks, gold[f"{wn}_tokens"], "tokens")
    _eq(lg, gold[f"{wn}_logits"], "logits")
    for k, v in acv.items():
        _eq(v, gold[f"{wn}_{k}"], k)


# ------------------------------------------------------------------------------------------- GPU
class GpuOps:
    """Adapter giving the C-ABI ops the same Python interface as oracle.CpuLib (so run_ops can drive it)."""

    def __init__(self):
        from tinyllama_cpp_b200 import capi
        self.c = capi
        capi.init(0)

    def encode_rows(self, x, dt):
        return self.c.write_rows(x, dt)

    def decode_rows(self, raw, dt, n):
        return self.c.read_rows(raw, dt, n)

    def read_row(self, raw, dt, n):
        return self.c.read_rows(np.ascontiguousarray(raw).reshape(1, -1), dt, n)[0]

    def rms_norm(self, x, dt, n_ctx, n, w, start_pos=0):
        return self.c.rms_norm(x, dt, n_ctx, n, w, start_pos)

    def silu(self, x, dt, n_ctx, n, start_pos=0):
        return self.c.silu(x, dt, n_ctx, n, start_pos)

    def mul(self, a, b, dt, n_ctx, n, start_pos=0):
        return self.c.mul(a, b, dt, n_ctx, n, start_pos)

    def add(self, a, b, dt, n_ctx, n, start_pos=0):
        return self.c.add(a, b, dt, n_ctx, n, start_pos)

    def rotary_emb(self, x, dt, n_ctx, n, d_head, start_pos=0):
        return self.c.rotary_emb(x, dt, n_ctx, n, d_head, start_pos)

    def qkv_attn(self, q, k, v, dt, n_ctx, H, G, D, max_ctx, start_pos=0):
        return self.c.qkv_attn(q, k, v, dt, n_ctx, H, G, D, max_ctx, start_pos)

    def matmul_2d(self, x, xdt, n_ctx, k, w, wdt, n_out, odt, out_1d=False, start_pos=0):
        wt = self.c.Weight(w, wdt, n_out, k)
        return self.c.matmul_2d(x, xdt, n_ctx, wt, odt, out_1d, start_pos)

    def token_embed(self, w, wdt, n_vocab, n_embd, tokens, odt, start_pos=0):
        wt = self.c.Weight(w, wdt, n_vocab, n_embd)
        return self.c.token_embed(wt, tokens, odt, start_pos)


@pytest.mark.gpu
def test_gpu_matches_golden_ops():
    gold = np.load(GOLD / "ops.npz")
    g = GpuOps()
    got = MG.run_ops(g)
    for k in gold.files:
        if k.startswith("wdeq_"):
            continue
        a, b = got[k], gold[k]
        if k.startswith("attn_") and not k.startswith("attn_last"):
            pass
        _eq(a, b, k)
    # dequantised weights straight from the device layout
    I = MG.op_inputs()
    for wn, wdt in WDT.items():
        w = g.c.Weight(W.quantize_payload(I["w_small"], wdt), wdt, 64, 2048)
        _eq(w.dequant(), gold[f"wdeq_{wn}"], f"wdeq_{wn}")


@pytest.mark.gpu
@pytest.mark.parametrize("wn", ["q4", "q8", "f16"])
def test_gpu_matches_golden_mini(wn):
    from tinyllama_cpp_b200 import capi
    gold = np.load(GOLD / "mini.npz")
    M = MG.MINI
    cfg = W.mini_config(n_layers=M["n_layers"], n_vocab=M["n_vocab"])
    e = capi.Engine(cfg, M["max_ctx"], WDT[wn]).load(W.synth_weights(cfg, WDT[wn], seed=M["seed"]))
    prompt = W.synth_prompt(M["prompt_seed"], M["n_prompt"], cfg.n_vocab)
    toks = e.generate(prompt, M["n_new"])
    _eq(toks, gold[f"{wn}_tokens"], "tokens")
    # logits of every step, teacher-forced on the golden tokens through the logits() entry point
    gl = gold[f"{wn}_logits"]
    all_toks = gold[f"{wn}_tokens"]
    lg0 = e.logits(all_toks[: M["n_prompt"]], 0)
    _eq(lg0, gl[0], "prefill logits")
    for i in range(1, M["n_new"]):
        n = M["n_prompt"] + i
        _eq(e.logits(all_toks[:n], n - 1), gl[i], f"logits step {i}")
    e.close()


def _full_case(name):
    from tinyllama_cpp_b200 import capi
    gold = np.load(GOLD / f"{name}.npz")
    wdt, n_prompt, n_new, max_ctx = MG.FULL[name]
    cfg = W.TINYLLAMA
    e = capi.Engine(cfg, max_ctx, wdt).load(W.synth_weights(cfg, wdt, seed=1))
    prompt = W.synth_prompt(7, n_prompt, cfg.n_vocab)
    toks = e.generate(prompt, n_new)
    gt = gold["tokens"]
    first_bad = next((i for i in range(toks.size) if toks[i] != gt[i]), None)
    assert first_bad is None, f"{name}: greedy tokens diverge at index {first_bad} (step {first_bad - n_prompt}); margin there {gold['margin'][max(first_bad - n_prompt, 0)]}"
    # logits of EVERY kept step, teacher forced on the golden tokens: the prefill call for step 0, then single rows on top of the
    # cache the greedy run left behind (rows < n - 1 are those of the golden sequence, which the run reproduced)
    for step, glog in zip(gold["keep_steps"], gold["keep_logits"]):
        step = int(step)
        n = n_prompt + step
        lg = e.logits(gt[:n], 0) if step == 0 else e.logits(gt[:n], n - 1)
        _eq(lg, glog, f"{name} logits step {step}")
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["full_q8", "full_f16", "full_q4"])
def test_gpu_full_model_greedy_identity(name):
    """BASELINE.json configs 1-3 at full size: greedy tokens identical to the reference CPU path."""
    if not (GOLD / f"{name}.npz").exists():
        pytest.skip(f"{name}.npz not generated")
    _full_case(name)
