"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes), against the CPU checker on the
same seeded inputs.  The checker is the real reference (oracle/_ref) when its prebuilt library travelled to
this box, else the plain-C restatement.  Everything is integer/bit exact: Q8 codes, fp16 bits, fp32 logits.
"""
import numpy as np
import pytest

import oracle
from oracle import F16, F32, Q4, Q8
from tinyllama_cpp_b200 import weights as W

pytestmark = pytest.mark.gpu
ADT = {F16: F16, Q8: Q8, Q4: Q8}


@pytest.fixture(scope="module")
def capi():
    from tinyllama_cpp_b200 import capi
    capi.init(0)
    return capi


def rand_rows(rng, rows, n, scale=1.0):
    return (rng.standard_normal((rows, n)) * scale).astype(np.float32)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_device_is_blackwell(capi):
    info = capi.device_info()
    assert info["cc"][0] == 10, info


@pytest.mark.parametrize("wdt", [Q4, Q8, F16])
@pytest.mark.parametrize("shape", [(8, 2048), (37, 5632), (256, 64)])
def test_dequantised_weights_bit_exact(capi, checker, wdt, shape):
    rows, cols = shape
    rng = np.random.default_rng(rows)
    pay = W.quantize_payload(rand_rows(rng, rows, cols, 0.02), wdt)
    w = capi.Weight(pay, wdt, rows, cols)
    ref = np.stack([checker.read_row(r, wdt, cols) for r in pay.reshape(rows, -1)])
    assert np.array_equal(bits(w.dequant()), bits(ref))
    assert w.nbytes() == pay.size          # repack keeps the byte count (roofline denominator unchanged)


@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 100, 2048, 5632])
def test_row_codecs(capi, checker, n):
    rng = np.random.default_rng(n)
    x = np.concatenate([rand_rows(rng, 2, n, s) for s in (0.0, 1e-4, 1.0, 40.0)])
    for dt in (Q8, F16, F32):
        enc = capi.write_rows(x, dt)
        assert np.array_equal(enc, checker.encode_rows(x, dt)), (n, dt)
        assert np.array_equal(bits(capi.read_rows(enc, dt, n)), bits(checker.decode_rows(enc, dt, n)))


def test_q8_rounding_ties(capi, checker):
    """roundf is half-away-from-zero; build rows whose scaled values sit exactly on .5 (quants.h:64)."""
    x = np.zeros((4, 32), np.float32)
    x[:, 0] = 127.0
    x[0, 1:] = np.arange(31) + 0.5
    x[1, 1:] = -(np.arange(31) + 0.5)
    x[2, 1:] = np.nextafter(np.float32(np.arange(31) + 0.5), np.float32(0))
    x[3, 1:] = np.nextafter(np.float32(np.arange(31) + 0.5), np.float32(1e9))
    assert np.array_equal(capi.write_rows(x, Q8), checker.encode_rows(x, Q8))


@pytest.mark.parametrize("wdt", [Q4, Q8, F16])
@pytest.mark.parametrize("shape", [(2048, 2048), (256, 2048), (96, 5632), (515, 2048)])
def test_matmul_2d(capi, checker, wdt, shape):
    """N = 515 (not a multiple of 32 or of the 4-row warp pass) is checked through the fp32 logits form only:
    the reference's block-aware row stride floors (tensor.h:97-117), so its encoded rows overlap for such N."""
    n_out, k = shape
    adt = ADT[wdt]
    rng = np.random.default_rng(n_out * 7 + wdt)
    n_ctx = 3
    x = checker.encode_rows(rand_rows(rng, n_ctx, k, 1.3), adt)
    pay = W.quantize_payload(rand_rows(rng, n_out, k, 0.02), wdt)
    w = capi.Weight(pay, wdt, n_out, k)
    for start in ((0, 2) if n_out % 32 == 0 else ()):
        got = capi.matmul_2d(x, adt, n_ctx, w, adt, start_pos=start)
        ref = checker.matmul_2d(x, adt, n_ctx, k, pay, wdt, n_out, adt, start_pos=start)
        assert np.array_equal(got[start:], ref[start:]), (wdt, shape, start)
    got = capi.matmul_2d(x, adt, n_ctx, w, F32, out_1d=True, start_pos=n_ctx - 1)
    ref = checker.matmul_2d(x, adt, n_ctx, k, pay, wdt, n_out, F32, out_1d=True, start_pos=n_ctx - 1)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("adt", [Q8, F16])
def test_elementwise_ops(capi, checker, adt):
    rng = np.random.default_rng(70 + adt)
    for n_ctx, n in ((4, 2048), (2, 5632), (3, 256)):
        x = checker.encode_rows(rand_rows(rng, n_ctx, n, 2.0), adt)
        y = checker.encode_rows(rand_rows(rng, n_ctx, n, 0.7), adt)
        wn = (1 + 0.1 * rng.standard_normal(n)).astype(np.float16)
        assert np.array_equal(capi.rms_norm(x, adt, n_ctx, n, wn), checker.rms_norm(x, adt, n_ctx, n, wn))
        assert np.array_equal(capi.silu(x, adt, n_ctx, n), checker.silu(x, adt, n_ctx, n))
        assert np.array_equal(capi.mul(x, y, adt, n_ctx, n), checker.mul(x, y, adt, n_ctx, n))
        assert np.array_equal(capi.add(x, y, adt, n_ctx, n), checker.add(x, y, adt, n_ctx, n))
        assert np.array_equal(capi.rotary_emb(x, adt, n_ctx, n, 64), checker.rotary_emb(x, adt, n_ctx, n, 64))
        a = capi.rotary_emb(x, adt, n_ctx, n, 64, start_pos=n_ctx - 1)
        assert np.array_equal(a, checker.rotary_emb(x, adt, n_ctx, n, 64, start_pos=n_ctx - 1))


def test_silu_extremes(capi, checker):
    """expf path: large |x| (under/overflow branches of glibc expf), zeros, tiny values."""
    v = np.array([0, -0.0, 1e-30, -1e-30, 20, -20, 87.9, -87.9, 88.5, -88.5, 89, -89, 103.9, -103.9, 104.5, -104.5, 1e4, -1e4,
                  0.5, -0.5, 3.25, -3.25, 7, -7, 15.5, -15.5, 31, -31, 50, -50, 65000, -65000], np.float32).reshape(1, 32)
    x = checker.encode_rows(v, F16)
    assert np.array_equal(capi.silu(x, F16, 1, 32), checker.silu(x, F16, 1, 32))
    x32 = np.ascontiguousarray(v).view(np.uint8).reshape(1, -1)
    assert np.array_equal(capi.silu(x32, F32, 1, 32), checker.silu(x32, F32, 1, 32))


def test_rms_norm_exact_sum_adversarial(capi, checker):
    """The in-order 2048-term fp32 sum (ops.h:765-767) is reproduced by a parallel exact algorithm; stress it."""
    rng = np.random.default_rng(5)
    n = 2048
    rows = []
    for mode in range(8):
        for _ in range(6):
            if mode == 0:
                r = rng.standard_normal(n)
            elif mode == 1:
                r = np.ldexp(rng.integers(1, 5, n).astype(np.float64), -rng.integers(0, 6, n))      # many rounding ties
            elif mode == 2:
                r = np.where(rng.random(n) < 0.3, 0.0, np.abs(rng.standard_normal(n)) * np.ldexp(1.0, rng.integers(-20, 20, n)))
            elif mode == 3:
                r = np.full(n, 1.0)                                                              # sum crosses many binades on exact powers of two
            elif mode == 4:
                r = np.concatenate([np.full(1, 300.0), np.full(n - 1, 1e-3)])                    # tiny terms under a big head
            elif mode == 5:
                r = np.concatenate([np.full(n - 1, 1e-3), np.full(1, 300.0)])
            elif mode == 6:
                r = np.zeros(n); r[rng.integers(0, n, 5)] = rng.standard_normal(5) * 10
            else:
                r = rng.standard_normal(n) * np.ldexp(1.0, rng.integers(-8, 8))
            rows.append(r.astype(np.float32))
    x = np.stack(rows)
    wn = np.ones(n, np.float16)
    x32 = np.ascontiguousarray(x).view(np.uint8).reshape(len(rows), -1)
    assert np.array_equal(capi.rms_norm(x32, F32, len(rows), n, wn), checker.rms_norm(x32, F32, len(rows), n, wn))
    for adt in (Q8, F16):
        # keep fp16 finite: inf activations give NaNs whose sign differs between x86 and CUDA (not reproduced)
        xe = checker.encode_rows(np.clip(x, -6e4, 6e4), adt)
        got, ref = capi.rms_norm(xe, adt, len(rows), n, wn), checker.rms_norm(xe, adt, len(rows), n, wn)
        bad = [i for i in range(len(rows)) if not np.array_equal(got[i], ref[i])]
        assert not bad, (adt, bad)


def test_megakernel_exact_sum_adversarial(capi):
    """The persistent kernel's own parallel exact sum against numpy's sequential float32 accumulation."""
    rng = np.random.default_rng(11)
    cases = []
    for n in (1, 3, 4, 5, 31, 64, 129, 300, 1000, 2047, 2048, 2049, 4096, 5000):
        cases.append(np.abs(rng.standard_normal(n)).astype(np.float32) ** 2)
        cases.append(np.exp(rng.standard_normal(n) * 3 - 4).astype(np.float32))
    n = 2048
    for _ in range(6):
        cases.append(np.ldexp(rng.integers(1, 5, n).astype(np.float64), -rng.integers(0, 6, n)).astype(np.float32))       # ties
        cases.append(np.where(rng.random(n) < 0.3, 0.0, np.abs(rng.standard_normal(n)) * np.ldexp(1.0, rng.integers(-20, 20, n))).astype(np.float32))
        cases.append((np.abs(rng.standard_normal(n)) * np.ldexp(1.0, rng.integers(-8, 8))).astype(np.float32))
        q = rng.integers(-127, 128, n).astype(np.float32) * np.float32(np.float16(rng.random() * 0.05))
        cases.append((q * q).astype(np.float32))                                                                           # squares of Q8 values
    cases += [np.full(n, 1.0, np.float32), np.concatenate([np.full(1, 300.0), np.full(n - 1, 1e-3)]).astype(np.float32),
              np.concatenate([np.full(n - 1, 1e-3), np.full(1, 300.0)]).astype(np.float32), np.zeros(n, np.float32),
              np.full(n, 1e-40, np.float32), np.full(n, 3e38 / n, np.float32)]
    z = np.zeros(n, np.float32); z[rng.integers(0, n, 5)] = np.abs(rng.standard_normal(5)) * 10
    cases.append(z)
    for i, t in enumerate(cases):
        want = np.float32(0.0)
        for v in t:
            want = np.float32(want + v)
        got = capi.selftest_exact_sum(t)
        assert got.view(np.uint32) == want.view(np.uint32), (i, t.size, got, want)


@pytest.mark.parametrize("adt,bdt,n", [(Q8, Q8, 64), (Q8, Q8, 2048), (Q8, Q4, 2048), (Q8, Q4, 5632), (F16, F16, 2048), (F16, F16, 67),
                                       (F32, F32, 2048), (F32, F32, 131), (F32, F32, 5)])
def test_vec_dot_product(capi, checker, adt, bdt, n):
    """ops::vec_dot_product (gten/ops.h:482-512) on host rows, every dtype pair of the reference's switch: the float returned
    is bit-identical to the reference's AVX build (4 integer lanes per block / 8 float lanes + in-order tail)."""
    rng = np.random.default_rng(n + 7 * adt + bdt)
    x, y = rand_rows(rng, 1, n, 1.3)[0], rand_rows(rng, 1, n, 0.02)[0]
    a = x if adt == F32 else checker.write_row(x, adt)
    b = y if bdt == F32 else (W.quantize_payload(y.reshape(1, -1), Q4) if bdt == Q4 else checker.write_row(y, bdt))
    got = capi.vec_dot_product(a, adt, b, bdt, n)
    want = checker.vec_dot(a, adt, b, bdt, n)
    assert bits(np.float32(got)) == bits(np.float32(want)), (got, want)


@pytest.mark.parametrize("wdt", [Q4, Q8, F16])
def test_token_embed(capi, checker, wdt):
    rng = np.random.default_rng(9)
    n_vocab, n = 50, 2048
    adt = ADT[wdt]
    pay = W.quantize_payload(rand_rows(rng, n_vocab, n, 0.02), wdt)
    toks = rng.integers(0, n_vocab, 6).astype(np.int32)
    w = capi.Weight(pay, wdt, n_vocab, n)
    assert np.array_equal(capi.token_embed(w, toks, adt), checker.token_embed(pay, wdt, n_vocab, n, toks, adt))
    got = capi.token_embed(w, toks, adt, start_pos=5)
    assert np.array_equal(got[5:], checker.token_embed(pay, wdt, n_vocab, n, toks, adt, start_pos=5)[5:])


@pytest.mark.parametrize("adt", [Q8, F16])
@pytest.mark.parametrize("n_ctx,max_ctx", [(1, 64), (7, 64), (8, 64), (9, 64), (33, 128), (40, 128), (64, 256), (100, 512)])
def test_attention(capi, checker, adt, n_ctx, max_ctx):
    """Multi-row and single-row calls inside the reference's valid domain."""
    assert (2 * n_ctx <= max_ctx) if adt == F16 else (((n_ctx + 31) // 32) * 34 <= max_ctx)
    rng = np.random.default_rng(n_ctx + adt)
    H, G, D = 32, 4, 64
    q = checker.encode_rows(rand_rows(rng, n_ctx, H * D), adt)
    k = checker.encode_rows(rand_rows(rng, n_ctx, G * D), adt)
    v = checker.encode_rows(rand_rows(rng, n_ctx, G * D), adt)
    for start in sorted({0, n_ctx - 1}):
        got = capi.qkv_attn(q, k, v, adt, n_ctx, H, G, D, max_ctx, start_pos=start)
        ref = checker.qkv_attn(q, k, v, adt, n_ctx, H, G, D, max_ctx, start_pos=start)
        assert np.array_equal(got[start:], ref[start:]), (adt, n_ctx, start)


@pytest.mark.parametrize("adt", [Q8, F16])
def test_attention_long_context_decode_row(capi, checker, adt):
    """A single decode row at t = 2048 (full KV, BASELINE config 3) against the reference itself (`_ref` where its library
    travelled).  max_ctx = 2176 keeps the Q8 score row of 64 blocks x 34 B inside its own stride."""
    rng = np.random.default_rng(2048)
    H, G, D, n_ctx, max_ctx = 32, 4, 64, 2048, 2176
    q = checker.encode_rows(rand_rows(rng, n_ctx, H * D), adt)
    k = checker.encode_rows(rand_rows(rng, n_ctx, G * D), adt)
    v = checker.encode_rows(rand_rows(rng, n_ctx, G * D), adt)
    got = capi.qkv_attn(q, k, v, adt, n_ctx, H, G, D, max_ctx, start_pos=n_ctx - 1)
    ref = checker.qkv_attn(q, k, v, adt, n_ctx, H, G, D, max_ctx, start_pos=n_ctx - 1)
    assert np.array_equal(got[-1], ref[-1])


@pytest.mark.parametrize("wdt", [Q4, Q8, F16])
def test_engine_mini_model(capi, checker, wdt):
    """Teacher-forced logits, free-running greedy tokens and every per-layer activation, bit for bit."""
    cfg = W.mini_config(n_layers=3, n_vocab=300)
    wl = list(W.synth_weights(cfg, wdt, seed=21))
    max_ctx = 128
    cm = checker.model(cfg, max_ctx, wdt).load(wl)
    e = capi.Engine(cfg, max_ctx, wdt).load(wl)
    prompt = W.synth_prompt(5, 33, cfg.n_vocab)
    n_new = 20
    ct, _, clog = cm.generate(prompt, n_new, want_logits=True)
    gt = e.generate(prompt, n_new)
    assert np.array_equal(gt, ct)
    # the default path is the persistent megakernel; the one-kernel-per-phase path (graph replay and eager) agrees
    e.set_option("mega", 0)
    assert np.array_equal(e.generate(prompt, n_new), ct)
    e.set_option("graph", 0)
    assert np.array_equal(e.generate(prompt, n_new), ct)
    e.set_option("mega", 1)
    # an EOS stops the device-side loop: generation ends with the first occurrence of that id
    eos = int(ct[33 + 5])
    first = 33 + int(np.argmax(ct[33:] == eos))
    got = e.generate(prompt, n_new, eos_id=eos)
    assert np.array_equal(got, ct[:first + 1])
    # logits() with the reference's calling protocol (all tokens so far + start_pos)
    assert np.array_equal(bits(e.logits(ct[:33], 0)), bits(clog[0]))
    for i in (1, 2, 7, n_new - 1):
        n = 33 + i
        assert np.array_equal(bits(e.logits(ct[:n], n - 1)), bits(clog[i])), i
    # per-layer activations of the last processed row
    e.set_option("capture_acv", 1)
    n = 33 + n_new - 1
    e.logits(ct[:n], n - 1)
    cm.capture_row(n - 1)
    cm.logits(ct[:n], n - 1)
    for layer in range(cfg.n_layers):
        for name, aid in oracle.LAYER_ACVS.items():
            assert np.array_equal(bits(e.acv(layer, aid)), bits(cm.acv(layer, aid, n - 1))), (layer, name)
    assert np.array_equal(bits(e.acv(0, oracle.A_FINAL_NORM)), bits(cm.acv(0, oracle.A_FINAL_NORM, n - 1)))
    assert np.array_equal(bits(e.acv(0, oracle.A_EMB)), bits(cm.acv(0, oracle.A_EMB, n - 1)))
    e.close(); cm.close()


def test_engine_errors(capi):
    cfg = W.mini_config(n_layers=1, n_vocab=64)
    e = capi.Engine(cfg, 16, Q4)
    with pytest.raises(capi.GtbError, match="not loaded"):
        e.logits(np.zeros(4, np.int32), 0)
    with pytest.raises(capi.GtbError, match="does not match the expected size"):
        e.set_weight(0, W.T_Q, np.zeros(10, np.uint8))
    e.load(W.synth_weights(cfg, Q4, seed=2))
    with pytest.raises(capi.GtbError, match="exceed provided maximum ctx size"):
        e.logits(np.zeros(17, np.int32), 0)
    e.close()


def test_gten_file_roundtrip(capi, checker, tmp_path):
    cfg = W.mini_config(n_layers=1, n_vocab=64)
    wl = list(W.synth_weights(cfg, Q8, seed=4))
    path = tmp_path / "mini.q8.gten"
    W.write_gten(path, cfg, Q8, wl)
    assert all(np.array_equal(a[2], b[2]) for a, b in zip(wl, W.read_gten(path, cfg, Q8)))
    e = capi.Engine(cfg, 32, Q8).load_gten(path)
    cm = checker.model(cfg, 32, Q8).load(wl)
    prompt = W.synth_prompt(1, 5, cfg.n_vocab)
    assert np.array_equal(e.generate(prompt, 6), cm.generate(prompt, 6)[0])
    e.close(); cm.close()


@pytest.mark.parametrize("wdt", [Q4, F16])
def test_gten_loader_pipeline_and_name_check(capi, checker, tmp_path, wdt):
    """gtb_engine_load_gten (mmap + pinned double-buffered staging + on-device repack): payloads larger than one 16 MB staging chunk,
    three layers; a file whose records are out of order, renamed or truncated is refused with the reference's kind of message."""
    cfg = W.mini_config(n_layers=3, n_vocab=9000)           # embedding / lm_head payloads of 10-37 MB: several staging chunks
    wl = list(W.synth_weights(cfg, wdt, seed=14))
    path = tmp_path / "m.gten"
    W.write_gten(path, cfg, wdt, wl)
    e = capi.Engine(cfg, 32, wdt).load_gten(path)
    cm = checker.model(cfg, 32, wdt).load(wl)
    prompt = W.synth_prompt(2, 6, cfg.n_vocab)
    assert np.array_equal(e.generate(prompt, 5), cm.generate(prompt, 5)[0])
    e.close(); cm.close()
    raw = bytearray(path.read_bytes())
    # rename the second record (q_proj of layer 0 -> k_proj): same length, wrong tensor
    i = raw.find(b"model.layers.0.self_attn.q_proj.weight")
    bad = bytearray(raw); bad[i:i + 38] = b"model.layers.0.self_attn.k_proj.weight"
    (tmp_path / "bad_name.gten").write_bytes(bad)
    (tmp_path / "short.gten").write_bytes(raw[: len(raw) // 2])
    badmagic = bytearray(raw); badmagic[0] ^= 0xff
    (tmp_path / "magic.gten").write_bytes(badmagic)
    for name, msg in (("bad_name.gten", "unexpected record"), ("short.gten", "truncated"), ("magic.gten", "Magic number")):
        f = capi.Engine(cfg, 32, wdt)
        with pytest.raises(capi.GtbError, match=msg):
            f.load_gten(tmp_path / name)
        f.close()
    # a file of another shape (different n_ffn) fails the reference's per-tensor size check (tinyllama.cpp:316-319)
    g = capi.Engine(W.mini_config(n_layers=3, n_vocab=9000, n_ffn=2816), 32, wdt)
    with pytest.raises(capi.GtbError, match="does not match the expected size"):
        g.load_gten(path)
    g.close()


def test_real_weights_golden_tokens(capi):
    """Optional: with GTEN_REAL_MODEL=/path/to/tinyllama.<fp16|q8|q4>.gten (the reference's own model files, README.md:6) the prompt
    ids and the first 30 greedy output ids must be those of the reference's golden comment (tinyllama.cpp:101-104)."""
    import os
    path = os.environ.get("GTEN_REAL_MODEL")
    if not path or not os.path.exists(path):
        pytest.skip("GTEN_REAL_MODEL not set (no network in this environment: the real checkpoint is not available)")
    wdt = {"fp16": F16, "q8": Q8, "q4": Q4}[[k for k in ("fp16", "q8", "q4") if f".{k}." in os.path.basename(path)][0]]
    prompt = np.array([1, 32001, 1404, 13, 22110, 338, 8425, 28579, 29973, 32002, 29871, 13, 32001, 20255, 13], np.int32)
    golden = [24115, 29880, 28579, 338, 263, 5332, 8578, 359, 13434, 322, 7766, 391, 1058, 338, 5545,
              697, 310, 278, 1556, 4100, 13994, 297, 278, 5849, 310, 28579, 391, 6368, 322, 6944]
    e = capi.Engine(W.TINYLLAMA, 64, wdt).load_gten(path)
    toks = e.generate(prompt, 30)
    e.close()
    assert toks[15:].tolist() == golden, toks[15:].tolist()


def test_config5_sequence_sample_matches_reference(capi, checker):
    """BASELINE.json configs[4] (64 independent Q4 sequences served as replicas): one of the job's sequences (the prompt
    recipe of bench.py --workload q4_seq64, sequence 5) at full size, greedy tokens identical to the CPU reference."""
    cfg = W.TINYLLAMA
    wl = list(W.synth_weights(cfg, Q4, seed=1))
    n_prompt, n_new = 128, 12
    prompt = W.synth_prompt(100 + 5, n_prompt, cfg.n_vocab)
    cm = checker.model(cfg, 160, Q4).load(wl)
    want = cm.generate(prompt, n_new)[0]
    e = capi.Engine(cfg, 160, Q4).load(wl)
    e.prefill(prompt)
    e.decode(n_new - 1)
    got = e.read_tokens(0, n_prompt + n_new)
    assert np.array_equal(got, want)
    e.close(); cm.close()


def test_device_topk_is_the_reference_candidate_set(capi):
    """gtb_engine_topk returns what topk_sample's partial_sort keeps (tinyllama.cpp:466-478): the k largest logits."""
    cfg = W.mini_config(n_layers=1, n_vocab=3000)
    e = capi.Engine(cfg, 32, Q4).load(W.synth_weights(cfg, Q4, seed=6))
    lg = e.logits(W.synth_prompt(4, 9, cfg.n_vocab), 0)
    for k in (1, 5, 40, 64):
        v, ids = e.topk(k)
        order = np.lexsort((np.arange(lg.size), -lg.astype(np.float64)))[:k]        # value descending, ties to the lower id
        assert np.array_equal(ids, order.astype(np.int32)), k
        assert np.array_equal(v.view(np.uint32), lg[order].view(np.uint32))
    with pytest.raises(capi.GtbError, match="argument check failed"):
        e.topk(65)
    e.close()
