"""The order-exact MULTI-ROW kernels (gtb_xrows.cu) against the reference itself (oracle `_ref` where its prebuilt library
travelled, else the pinned C port) and against the row-at-a-time kernels.

Contract: BIT equality.  R rows share every weight load but each (row, output, lane) runs the reference's own ordered chain
(gten/ops.h:282-292, 765-767, 982-988, 181-197), so
  * the exact prefill in passes of up to 64 rows gives the logits, K/V cache and greedy tokens of the reference's row loop
    (ops.h:632) -- every pass boundary and every P.V lane-split case (n % 8, t < n8) is exercised;
  * the exact batched decode gives, for every slot, the tokens and logits of the same sequence decoded alone by the reference
    (BASELINE.json configs[4] at full size: 4 sequences of the q4_seq64 recipe against committed golden tokens).
"""
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import F16, Q4, Q8
from tinyllama_cpp_b200 import weights as W

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(GOLD))
import make_golden as MG  # noqa: E402


@pytest.fixture(scope="module")
def capi():
    from tinyllama_cpp_b200 import capi
    capi.init(0)
    return capi


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("wdt", [Q8, Q4, F16], ids=["q8", "q4", "f16"])
@pytest.mark.parametrize("n_prompt", [4, 7, 16, 63, 64, 65, 100, 129, 190])
def test_exact_prefill_matches_reference(capi, checker, wdt, n_prompt):
    """Prefill logits (bit for bit) and the following greedy tokens equal the CPU reference's; the same engine with the
    multi-row path switched off (row-at-a-time persistent kernel) gives the same bits."""
    cfg = W.mini_config(n_layers=2, n_vocab=300)
    wl = list(W.synth_weights(cfg, wdt, seed=31))
    max_ctx = 400 if wdt == F16 else 256   # the oracle's prefill rows need ceil(n/32)*34 <= max_ctx (Q8) / 2n <= max_ctx (FP16), SURVEY App. B1
    n_new = 6
    cm = checker.model(cfg, max_ctx, wdt).load(wl)
    prompt = W.synth_prompt(9, n_prompt, cfg.n_vocab)
    want_toks, _, want_lg = cm.generate(prompt, n_new, want_logits=True)
    e = capi.Engine(cfg, max_ctx, wdt).load(wl)
    e.set_option("xr_rows", 64)        # 64 rows per pass: prompts of 65, 100, 129, 190 rows cross pass boundaries
    got = e.logits(prompt, 0)
    assert np.array_equal(bits(got), bits(want_lg[0])), f"prefill logits differ from the {checker.kind} oracle"
    e.set_option("xr_rows", 512)       # the default: one pass
    assert np.array_equal(bits(e.logits(prompt, 0)), bits(got))
    e.set_option("xr_rows", 24)        # small passes: the per-head attention kernel
    assert np.array_equal(bits(e.logits(prompt, 0)), bits(got))
    e.set_option("xr_rows", 64)
    toks = e.generate(prompt, n_new)
    assert np.array_equal(toks, want_toks), (toks[n_prompt:], want_toks[n_prompt:])
    assert np.array_equal(bits(e.read_logits()), bits(want_lg[-1]))
    e.set_option("xrows", 0)
    assert np.array_equal(bits(e.logits(prompt, 0)), bits(got))
    e.close()
    cm.close()


@pytest.mark.parametrize("wdt", [Q8, Q4, F16], ids=["q8", "q4", "f16"])
def test_exact_multi_row_continuation(capi, checker, wdt):
    """logits(tokens, start_pos) with several new rows on top of an existing cache (tinyllama.cpp:45-61 with start_pos > 0):
    the rows use the call's n_ctx for the P.V lane split."""
    cfg = W.mini_config(n_layers=2, n_vocab=300)
    wl = list(W.synth_weights(cfg, wdt, seed=32))
    cm = checker.model(cfg, 320, wdt).load(wl)
    e = capi.Engine(cfg, 320, wdt).load(wl)
    toks = W.synth_prompt(4, 150, cfg.n_vocab)
    for a, b in ((0, 40), (40, 53), (53, 54), (54, 150)):
        want = cm.logits(toks[:b], a)
        got = e.logits(toks[:b], a)
        assert np.array_equal(bits(got), bits(want)), (a, b)
    e.close()
    cm.close()


@pytest.mark.parametrize("graph", [1, 0], ids=["graph", "eager"])
@pytest.mark.parametrize("wdt,lens", [(Q4, (20, 37, 64)), (Q8, (5, 33, 40, 41, 64, 90, 100, 7)), (Q4, (50,)),
                                      (Q4, tuple(range(3, 3 + 37 * 3, 3))), (Q8, tuple(range(10, 100, 10))), (F16, (9, 33, 64, 70, 12))],
                         ids=["q4x3", "q8x8", "q4x1", "q4x37", "q8x9", "f16x5"])
def test_exact_batch_decode_matches_reference(capi, checker, wdt, lens, graph):
    """Every slot of the exact batched decode: tokens and last logits bit-identical to the SAME sequence generated alone by
    the CPU reference (different prompt lengths per slot: own positions, own K/V)."""
    cfg = W.mini_config(n_layers=2, n_vocab=300)
    wl = list(W.synth_weights(cfg, wdt, seed=6))
    steps = 7
    mc = 160
    e = capi.Engine(cfg, mc, wdt).load(wl)
    e.set_option("graph", graph)
    e.batch_create(len(lens))
    prompts = [W.synth_prompt(40 + i, n, cfg.n_vocab) for i, n in enumerate(lens)]
    for s, p in enumerate(prompts):
        if s % 2:
            e.batch_prefill(s, p)             # multi-row prefill straight into the slot
        else:
            e.prefill(p)                      # the engine's own sequence, then the slot takes over tokens, position and K/V
            e.batch_adopt(s)
    e.batch_decode(3)
    e.batch_decode(steps - 3)
    cm = checker.model(cfg, 160, wdt).load(wl)
    for s, p in enumerate(prompts):
        n = len(p)
        want, _, lg = cm.generate(p, steps + 1, want_logits=True)
        assert e.batch_position(s) == n + steps
        got = e.batch_read_tokens(s, 0, n + steps + 1)
        assert np.array_equal(got, want), (s, got[n:], want[n:])
        assert np.array_equal(bits(e.batch_read_logits(s)), bits(lg[-1])), s
    cm.close()
    e.close()


def test_exact_batch_equals_single_sequence_long_context(capi):
    """Contexts past several K/V tiles (128 / 64 positions) and past a multiple of 8: slots vs the row-at-a-time kernels."""
    cfg = W.mini_config(n_layers=2, n_vocab=300)
    wdt = Q4
    wl = list(W.synth_weights(cfg, wdt, seed=8))
    lens = (700, 333, 256, 257, 1)
    steps = 9
    e = capi.Engine(cfg, 720, wdt).load(wl)
    e.batch_create(len(lens))
    prompts = [W.synth_prompt(70 + i, n, cfg.n_vocab) for i, n in enumerate(lens)]
    for s, p in enumerate(prompts):
        e.batch_prefill(s, p)
    e.batch_decode(steps)
    one = capi.Engine(cfg, 720, wdt).load(wl)
    one.set_option("xrows", 0)                # every row through the persistent kernel
    for s, p in enumerate(prompts):
        one.prefill(p)
        one.decode(steps)
        n = len(p)
        assert np.array_equal(e.batch_read_tokens(s, 0, n + steps + 1), one.read_tokens(0, n + steps + 1)), s
        assert np.array_equal(bits(e.batch_read_logits(s)), bits(one.read_logits())), s
    one.close()
    e.close()


def test_exact_batch_eos_leave_and_rejoin(capi, checker):
    """Option "batch_eos": a slot that samples the id stops there (the reference's `break`, tinyllama.cpp:426) and leaves the batch; the
    other slots keep producing the reference's tokens; the freed slot takes a new prompt between decode calls (join)."""
    cfg = W.mini_config(n_layers=2, n_vocab=300)
    wl = list(W.synth_weights(cfg, Q4, seed=6))
    cm = checker.model(cfg, 160, Q4).load(wl)
    lens = (20, 37, 64, 9)
    steps = 12
    prompts = [W.synth_prompt(40 + i, n, cfg.n_vocab) for i, n in enumerate(lens)]
    want = [cm.generate(p, steps + 1)[0] for p in prompts]
    # an id that slot 1 samples in the middle of the run and that no other slot samples: that slot must stop, alone
    cands = [int(t) for t in want[1][lens[1] + 2: lens[1] + steps - 2]
             if all(int(t) not in want[s][lens[s]:] for s in (0, 2, 3)) and list(want[1][lens[1]:]).index(t) >= 2]
    assert cands, "no suitable EOS id in this synthetic run"
    eos = cands[0]
    stop_at = lens[1] + list(want[1][lens[1]:]).index(eos)          # index of the EOS token in slot 1's sequence
    e = capi.Engine(cfg, 160, Q4).load(wl)
    e.set_option("batch_eos", eos)
    e.batch_create(len(lens))
    for s, p in enumerate(prompts):
        e.batch_prefill(s, p)
    e.batch_decode(steps)
    for s, p in enumerate(prompts):
        if s == 1:
            assert e.batch_position(s) == stop_at
            assert np.array_equal(e.batch_read_tokens(s, 0, stop_at + 1), want[s][: stop_at + 1])
        else:
            assert e.batch_position(s) == len(p) + steps
            assert np.array_equal(e.batch_read_tokens(s, 0, len(p) + steps + 1), want[s]), s
    # join: the freed slot takes a new prompt; it and the others continue with the reference's tokens
    p_new = W.synth_prompt(77, 15, cfg.n_vocab)
    e.set_option("batch_eos", -1)
    e.batch_prefill(1, p_new)
    e.batch_decode(4)
    w_new = cm.generate(p_new, 5)[0]
    assert np.array_equal(e.batch_read_tokens(1, 0, 15 + 5), w_new)
    w0 = cm.generate(prompts[0], steps + 5)[0]
    assert np.array_equal(e.batch_read_tokens(0, 0, lens[0] + steps + 5), w0)
    e.close()
    cm.close()


def test_exact_batch_argument_errors(capi):
    cfg = W.mini_config(n_layers=1, n_vocab=64)
    e = capi.Engine(cfg, 32, Q4).load(W.synth_weights(cfg, Q4, seed=1))
    with pytest.raises(capi.GtbError):
        e.batch_create(65)
    e.batch_create(64)
    e.batch_create(2)
    with pytest.raises(capi.GtbError):
        e.batch_prefill(2, np.array([1, 2, 3], np.int32))
    e.batch_prefill(0, np.array([1, 2, 3], np.int32))
    e.batch_prefill(1, np.array([1, 2, 3], np.int32))
    with pytest.raises(capi.GtbError):
        e.batch_decode(40)                # past max_ctx
    e.batch_decode(2)
    assert e.batch_position(0) == e.batch_position(1) == 5
    assert np.array_equal(e.batch_read_tokens(0, 0, 6), e.batch_read_tokens(1, 0, 6))
    e.set_option("batch_exact", 0)
    with pytest.raises(capi.GtbError):
        e.batch_create(17)                # the order-free kernels take 16 sequences at most
    e.close()


def test_full_size_seq64_batch_identity(capi):
    """BASELINE.json configs[4] at full size: the first 4 sequences of bench.py's q4_seq64 recipe (128-token prompt + 64 new
    tokens, TinyLlama-1.1B Q4) decoded TOGETHER by the exact batched path -- tokens of every slot identical to the reference's
    own greedy run of that sequence, last logits bit-identical (tests/golden/seq64_q4.npz, generated from the unmodified
    reference by tests/golden/make_golden.py seq64_q4)."""
    f = GOLD / "seq64_q4.npz"
    if not f.exists():
        pytest.skip("seq64_q4.npz not generated")
    S = MG.SEQ64
    gold = np.load(f)
    cfg = W.TINYLLAMA
    e = capi.Engine(cfg, S["max_ctx"], Q4).load(W.synth_weights(cfg, Q4, seed=1))
    e.batch_create(S["n_seq"])
    for s in range(S["n_seq"]):
        e.batch_prefill(s, W.synth_prompt(100 + s, S["n_prompt"], cfg.n_vocab))
    e.batch_decode(S["n_new"] - 1)
    for s in range(S["n_seq"]):
        got = e.batch_read_tokens(s, 0, S["n_prompt"] + S["n_new"])
        want = gold["tokens"][s]
        bad = next((i for i in range(got.size) if got[i] != want[i]), None)
        assert bad is None, f"sequence {s}: tokens diverge at index {bad}"
        assert np.array_equal(bits(e.batch_read_logits(s)), bits(gold["last_logits"][s])), s
    e.close()


@pytest.mark.parametrize("wdt", [Q8, Q4], ids=["q8", "q4"])
def test_tensor_core_linears_are_bit_identical(capi, checker, wdt):
    """Option "xr_tensor": the multi-row Linears on tcgen05 (gtb_xtensor.cuh) -- integer lane sums out of kind::f16 MMAs, ordered fp32
    chains in the epilogue warps -- against the reference: prefill logits for pass shapes that select 2, 4 and 8 rows per thread,
    a continuation, and a batch of sequences at different positions."""
    cfg = W.mini_config(n_layers=2, n_vocab=300)
    wl = list(W.synth_weights(cfg, wdt, seed=33))
    cm = checker.model(cfg, 256, wdt).load(wl)
    e = capi.Engine(cfg, 256, wdt).load(wl)
    e.set_option("xr_tensor", 1)
    try:
        for n in (5, 37, 130, 190):
            prompt = W.synth_prompt(11, n, cfg.n_vocab)
            assert np.array_equal(bits(e.logits(prompt, 0)), bits(cm.logits(prompt, 0))), n
        toks = W.synth_prompt(4, 150, cfg.n_vocab)
        for a, b in ((0, 40), (40, 53), (53, 150)):
            assert np.array_equal(bits(e.logits(toks[:b], a)), bits(cm.logits(toks[:b], a))), (a, b)
        lens = (5, 33, 40, 41, 64, 90, 100, 7, 12)
        e.batch_create(len(lens))
        prompts = [W.synth_prompt(40 + i, n, cfg.n_vocab) for i, n in enumerate(lens)]
        for s, p in enumerate(prompts):
            e.batch_prefill(s, p)
        e.batch_decode(6)
        for s, p in enumerate(prompts):
            want, _, lg = cm.generate(p, 7, want_logits=True)
            assert np.array_equal(e.batch_read_tokens(s, 0, len(p) + 7), want), s
            assert np.array_equal(bits(e.batch_read_logits(s)), bits(lg[-1])), s
    finally:
        e.set_option("xr_tensor", 0)
        e.close()
        cm.close()


def test_full_size_f16_batch_identity(capi):
    """BASELINE.json configs[0] (FP16, 128-token prompt) at full size through the exact batched path: the slot that holds the golden
    prompt reproduces the reference's greedy tokens while another sequence advances in the same steps."""
    f = GOLD / "full_f16.npz"
    if not f.exists():
        pytest.skip("full_f16.npz not generated")
    gold = np.load(f)
    wdt, n_prompt, n_new, max_ctx = MG.FULL["full_f16"]
    cfg = W.TINYLLAMA
    steps = 48
    e = capi.Engine(cfg, max_ctx, wdt).load(W.synth_weights(cfg, wdt, seed=1))
    e.batch_create(2)
    e.batch_prefill(0, W.synth_prompt(7, n_prompt, cfg.n_vocab))
    e.batch_prefill(1, W.synth_prompt(8, 77, cfg.n_vocab))
    e.batch_decode(steps)
    got = e.batch_read_tokens(0, 0, n_prompt + steps + 1)
    want = gold["tokens"][: n_prompt + steps + 1]
    bad = next((i for i in range(got.size) if got[i] != want[i]), None)
    assert bad is None, f"tokens diverge at index {bad}"
    assert e.batch_position(1) == 77 + steps
    e.close()
