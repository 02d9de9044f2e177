"""Order-free ("fast") decode kernels (gtb_fastdec.cuh, option fast_decode) against the CPU checker.

Same contract as the batched prefill (tests/test_prefill_gpu.py): the kernels keep the reference's operations and re-encode
points (ops.h:645-646, 733-753, 762-804, 870-898, 996, 1084) but not its summation order, so parity is the STATED
TOLERANCE of DESIGN.md 4.4 -- SLACK x the distance between the reference's own two builds (scalar vs AVX,
tests/golden/order_sensitivity.npz) -- not bit equality.  The order-exact path stays the default and the bit-checked one.
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import Q4, Q8, F16
from tinyllama_cpp_b200 import weights as W

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
SLACK = 2.0
ABS = 5e-3
SENS = np.load(GOLD / "order_sensitivity.npz")


@pytest.fixture(scope="module")
def capi():
    from tinyllama_cpp_b200 import capi
    capi.init(0)
    return capi


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


MODES = {"persistent": {"fd_mega": 1}, "pdl-graph": {"fd_mega": 0, "graph": 1}, "pdl-eager": {"fd_mega": 0, "graph": 0}}


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("wdt", [Q8, Q4])
@pytest.mark.parametrize("n_prompt", [100, 64, 7])
def test_fast_rows_mini(capi, checker, wdt, n_prompt, mode):
    """The PDL-chained kernels (the default) and the persistent cooperative kernel (k_fd_mega) run the same phase code.
    Every row of the prompt through the fast kernels (row-at-a-time, like the reference's own prefill): logits of the
    last row within the reference's own build-to-build spread; then the next row through the ORDER-EXACT kernels on top of
    the K/V cache the fast kernels wrote (catches any cache-layout mistake)."""
    cfg = W.mini_config(n_layers=3, n_vocab=300)
    wl = list(W.synth_weights(cfg, wdt, seed=21))
    max_ctx = 192
    cm = checker.model(cfg, max_ctx, wdt).load(wl)
    e = capi.Engine(cfg, max_ctx, wdt).load(wl)
    e.set_option("fast_decode", 1)
    for k, v in MODES[mode].items():
        e.set_option(k, v)
    assert not e.uses_megakernel()
    prompt = W.synth_prompt(5, n_prompt, cfg.n_vocab)
    want = cm.logits(prompt, 0)
    got = e.logits(prompt, 0)
    wn = {Q8: "q8", Q4: "q4"}[wdt]
    sens = float(SENS[f"mini_{wn}_{n_prompt}_logits"])
    bound = SLACK * sens + min(ABS, 0.5 * sens)
    err = rel(got, want)
    print(f"\nfast rows wdt={wn} T={n_prompt}: logits rel {err:.2e} (reference scalar-vs-AVX {sens:.2e})")
    assert err <= bound, (err, sens)
    nxt = int(np.argmax(want))
    toks = np.concatenate([prompt, [nxt]]).astype(np.int32)
    e.set_option("fast_decode", 0)
    got_next = e.logits(toks, n_prompt)
    want_next = cm.logits(toks, n_prompt)
    assert rel(got_next, want_next) <= bound, rel(got_next, want_next)
    e.close(); cm.close()


@pytest.mark.parametrize("mega", [1, 0])
@pytest.mark.parametrize("wdt", [Q8, Q4])
def test_fast_single_row_on_exact_cache(capi, checker, wdt, mega):
    """One fast row on top of a K/V cache written by the exact path: a single row's worth of reordering noise only."""
    cfg = W.mini_config(n_layers=3, n_vocab=300)
    wl = list(W.synth_weights(cfg, wdt, seed=4))
    cm = checker.model(cfg, 192, wdt).load(wl)
    e = capi.Engine(cfg, 192, wdt).load(wl)
    prompt = W.synth_prompt(9, 150, cfg.n_vocab)
    want = cm.logits(prompt, 0)
    e.prefill(prompt[:-1])
    e.set_option("fast_decode", 1)
    e.set_option("fd_mega", mega)
    got = e.logits(prompt, prompt.size - 1)
    sens = max(float(SENS[k]) for k in SENS.files if k.startswith("mini_") and k.endswith("_logits"))
    assert rel(got, want) <= SLACK * sens + ABS, rel(got, want)
    e.close(); cm.close()


@pytest.mark.parametrize("mega", [1, 0])
def test_fast_greedy_loop_and_positions(capi, mega):
    """decode() under fast_decode: device-side greedy loop, positions advance, tokens are valid ids, and the sequence is
    reproducible (the kernels have no run-to-run nondeterminism: no float atomics)."""
    cfg = W.mini_config(n_layers=2, n_vocab=300)
    seqs = []
    for _ in range(2):
        e = capi.Engine(cfg, 256, Q4).load(W.synth_weights(cfg, Q4, seed=3))
        e.set_option("fast_decode", 1)
        e.set_option("fd_mega", mega)
        prompt = W.synth_prompt(2, 40, cfg.n_vocab)
        e.prefill(prompt)
        assert e.position() == 40
        e.decode(20)
        assert e.position() == 60
        toks = e.read_tokens(0, 61)
        assert np.array_equal(toks[:40], prompt)
        assert ((toks[40:] >= 0) & (toks[40:] < cfg.n_vocab)).all()
        seqs.append(toks)
        e.close()
    assert np.array_equal(seqs[0], seqs[1])


def test_fast_decode_refuses_activation_capture(capi):
    """capture_acv dumps the per-op activations of the order-exact kernels; the order-free chain keeps them in staged form only."""
    cfg = W.mini_config(n_layers=1, n_vocab=64)
    e = capi.Engine(cfg, 32, Q4).load(W.synth_weights(cfg, Q4, seed=1))
    e.set_option("fast_decode", 1)
    e.set_option("capture_acv", 1)
    with pytest.raises(capi.GtbError):
        e.logits(np.array([1, 2, 3], np.int32), 0)
    e.set_option("capture_acv", 0)
    assert e.logits(np.array([1, 2, 3], np.int32), 0).shape == (64,)
    e.close()


def test_fast_decode_rejects_fp16_models(capi):
    cfg = W.mini_config(n_layers=1, n_vocab=64)
    e = capi.Engine(cfg, 32, F16).load(W.synth_weights(cfg, F16, seed=1))
    e.set_option("fast_decode", 1)
    with pytest.raises(capi.GtbError):
        e.logits(np.array([1, 2, 3], np.int32), 0)
    e.close()


def test_fast_persistent_equals_pdl_chain(capi):
    """Both launch schemes run the same phase functions on the same decomposition (rows per CTA and rows per warp pass differ,
    the arithmetic of a row does not): bit-identical logits and tokens."""
    cfg = W.mini_config(n_layers=2, n_vocab=300)
    wl = list(W.synth_weights(cfg, Q4, seed=8))
    prompt = W.synth_prompt(4, 90, cfg.n_vocab)
    out, toks = [], []
    for mega in (1, 0):
        e = capi.Engine(cfg, 128, Q4).load(wl)
        e.prefill(prompt[:-1])
        e.set_option("fast_decode", 1)
        e.set_option("fd_mega", mega)
        out.append(e.logits(prompt, prompt.size - 1))
        e.decode(5)
        toks.append(e.read_tokens(0, prompt.size + 6))
        e.close()
    assert np.array_equal(out[0].view(np.uint32), out[1].view(np.uint32)), rel(out[0], out[1])
    assert np.array_equal(toks[0], toks[1])


def test_fast_generate_with_eos(capi):
    """generate() under fast_decode stops at the EOS id like the exact path (tinyllama.cpp:426)."""
    cfg = W.mini_config(n_layers=2, n_vocab=300)
    wl = list(W.synth_weights(cfg, Q4, seed=3))
    prompt = W.synth_prompt(2, 12, cfg.n_vocab)
    e = capi.Engine(cfg, 64, Q4).load(wl)
    e.set_option("fast_decode", 1)
    free = e.generate(prompt, 10)
    assert free.size == 22
    eos = int(free[12 + 4])
    first = 12 + int(np.argmax(free[12:] == eos))
    stopped = e.generate(prompt, 10, eos_id=eos)
    assert np.array_equal(stopped, free[: first + 1]), (stopped, free)
    e.close()


# ---------------------------------------------------------------- batched decode (gtb_engine_batch_*)
@pytest.mark.parametrize("graph", [1, 0], ids=["graph", "eager"])
@pytest.mark.parametrize("wdt,lens", [(Q4, (20, 37, 64)), (Q8, (5, 33, 40, 41, 64, 90, 100, 7)), (Q4, (50,)),
                                      (Q4, (3, 9, 17, 20, 31, 32, 33, 40, 47, 55, 64, 65, 70, 90, 100, 110)), (Q8, tuple(range(10, 100, 10)))],
                         ids=["q4x3", "q8x8", "q4x1", "q4x16", "q8x9"])
def test_batch_decode_equals_single_sequence(capi, wdt, lens, graph):
    """A sequence decoded in a batch gives the SAME BITS (tokens and logits) as the same sequence decoded alone through the
    order-free kernels: the batch only shares the weight reads.  Sequences have different lengths (own positions, own K/V)."""
    cfg = W.mini_config(n_layers=2, n_vocab=300)
    wl = list(W.synth_weights(cfg, wdt, seed=6))
    steps = 7
    e = capi.Engine(cfg, 128, wdt).load(wl)
    e.set_option("graph", graph)
    e.set_option("batch_exact", 0)        # the order-free batch kernels (the default batch mode is the exact one, tests/test_xrows_gpu.py)
    e.batch_create(len(lens))
    prompts = [W.synth_prompt(40 + i, n, cfg.n_vocab) for i, n in enumerate(lens)]
    for s, p in enumerate(prompts):
        e.prefill(p)                      # exact path; slot s takes over tokens, position and K/V cache
        e.batch_adopt(s)
    e.batch_decode(steps)
    for s, p in enumerate(prompts):
        one = capi.Engine(cfg, 128, wdt).load(wl)
        one.prefill(p)
        one.set_option("fast_decode", 1)
        one.set_option("fd_chunk", 128)       # the batch's attention chunk length (the single-sequence default is 64: lower latency)
        one.decode(steps)
        n = len(p)
        assert e.batch_position(s) == one.position() == n + steps
        want = one.read_tokens(0, n + steps + 1)
        got = e.batch_read_tokens(s, 0, n + steps + 1)
        assert np.array_equal(got, want), (s, got[n:], want[n:])
        assert np.array_equal(e.batch_read_logits(s).view(np.uint32), one.read_logits().view(np.uint32)), s
        one.close()
    e.close()


def test_batch_decode_argument_errors(capi):
    cfg = W.mini_config(n_layers=1, n_vocab=64)
    e = capi.Engine(cfg, 32, Q4).load(W.synth_weights(cfg, Q4, seed=1))
    e.set_option("batch_exact", 0)
    with pytest.raises(capi.GtbError):
        e.batch_create(17)
    with pytest.raises(capi.GtbError):
        e.batch_decode(1)                 # no slots
    e.batch_create(2)
    with pytest.raises(capi.GtbError):
        e.batch_adopt(2)
    e.prefill(np.array([1, 2, 3], np.int32))
    e.batch_adopt(0); e.batch_adopt(1)
    with pytest.raises(capi.GtbError):
        e.batch_decode(40)                # past max_ctx
    e.batch_decode(2)
    assert e.batch_position(0) == e.batch_position(1) == 5
    assert np.array_equal(e.batch_read_tokens(0, 0, 6), e.batch_read_tokens(1, 0, 6))
    e.batch_create(0)
    e.close()
    f = capi.Engine(cfg, 32, F16).load(W.synth_weights(cfg, F16, seed=1))
    f.set_option("batch_exact", 0)        # FP16 models batch through the exact multi-row kernels only
    with pytest.raises(capi.GtbError):
        f.batch_create(2)
    f.close()


# ---------------------------------------------------------------- token-level contract at full size
MARGIN_BOUND = 0.35       # logits; the reference's median top-1/top-2 margin on these weights is 0.15, sigma(logit) 0.90


@pytest.mark.parametrize("name,steps", [("full_q4", 512), ("full_q8", 256)])
def test_fast_decode_token_agreement_full_size(capi, name, steps):
    """What the tolerance means in tokens.  Teacher-forced on the reference's own greedy sequence at full size (golden tokens of
    BASELINE.json configs[2] / [1], generated from the unmodified reference), every step's row runs through the order-free
    kernels on an exact K/V cache.  Measured: top-1 agreement 80 % (Q4, 512 steps) / 83 % (Q8, 256 steps); EVERY disagreement
    sits at a step where the reference's own top-1/top-2 margin is below 0.26 logits (stated bound: 0.35), i.e. the order-free
    path only flips near-ties -- which with random-init weights is one step in five.  It is NOT greedy-identical (the exact
    kernels are), and this test pins how far from it it is."""
    import sys
    from pathlib import Path
    if not (GOLD / f"{name}.npz").exists():
        pytest.skip(f"{name}.npz not generated")
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    from fastdec_agreement import agreement
    r = agreement(name, steps)
    print(f"{name}: agreement {r['agree']}/{r['steps']} = {r['agreement_rate']:.3f}, largest reference margin at a mismatch "
          f"{r['max_mismatch_margin']:.3f}, logits rel-L2 mean {r['mean_rel_l2']:.3f} max {r['max_rel_l2']:.3f}")
    assert r["agreement_rate"] >= 0.70, r["agreement_rate"]
    assert r["max_mismatch_margin"] <= MARGIN_BOUND, [m for m in r["mismatches"] if m["ref_margin"] > MARGIN_BOUND]
    assert r["max_rel_l2"] <= 0.15, r["max_rel_l2"]
