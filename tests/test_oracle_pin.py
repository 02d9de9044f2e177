"""Pins oracle/gten_oracle.c (the plain-C restatement) bit-for-bit against oracle/_ref (the unmodified
reference, compiled from /root/reference with its own -O3 -fopenmp -mavx -mf16c line).

The reference ships no tests or golden vectors for this path, so the pin is the
reference itself run here; tests/test_golden.py additionally checks both against committed vectors.
"""
import numpy as np
import pytest

import oracle
from oracle import F16, F32, Q4, Q8
from tinyllama_cpp_b200 import weights as W

ADTS = [Q8, F16]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def rand_rows(rng, rows, n, scale=1.0):
    return (rng.standard_normal((rows, n)) * scale).astype(np.float32)


def test_fp16_widening_exhaustive(port, ref):
    for h in range(65536):
        a, b = port.fp16_to_fp32(h), ref.fp16_to_fp32(h)
        assert a.view(np.uint32) == b.view(np.uint32) or (np.isnan(a) and np.isnan(b)), hex(h)


def test_fp16_narrowing(port, ref):
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.standard_normal(4000).astype(np.float32) * s for s in (1e-8, 1e-6, 1e-4, 1e-2, 1, 300, 70000)])
    # every half value, its neighbours and the midpoints between consecutive halves (RNE ties)
    hv = np.arange(0, 0x7C00, dtype=np.uint16).view(np.float16).astype(np.float32)
    mids = ((hv[:-1].astype(np.float64) + hv[1:].astype(np.float64)) / 2).astype(np.float32)
    special = np.array([0, -0.0, np.inf, -np.inf, np.nan, 65504, 65519.996, 65520, 65536, 1e10, 2.9802322e-8, 2.98e-8, 3e-8], np.float32)
    xs = np.concatenate([xs, hv, -hv, mids, -mids, np.nextafter(mids, np.float32(0)), np.nextafter(mids, np.float32(1e9)), special])
    for x in xs:
        assert port.fp32_to_fp16(x) == ref.fp32_to_fp16(x), float(x)


@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 64, 100, 2048, 5632])
def test_row_codecs(port, ref, n):
    rng = np.random.default_rng(n)
    for scale in (0.0, 1e-4, 1.0, 40.0):
        x = rand_rows(rng, 1, n, scale)[0]
        for dt in (Q8, F16, F32):
            a, b = port.write_row(x, dt), ref.write_row(x, dt)
            assert np.array_equal(a, b)
            assert np.array_equal(bits(port.read_row(a, dt, n)), bits(ref.read_row(a, dt, n)))


def test_q4_dequant(port, ref):
    rng = np.random.default_rng(3)
    w = rand_rows(rng, 4, 2048, 0.02)
    pay = W.quantize_payload(w, Q4).reshape(4, -1)
    for r in pay:
        a, b = port.read_row(r, Q4, 2048), ref.read_row(r, Q4, 2048)
        assert np.array_equal(bits(a), bits(b))
    # numpy restatement of the decoder agrees too
    assert np.array_equal(bits(W.dequantize_payload(pay, Q4, 4, 2048)), bits(np.stack([ref.read_row(r, Q4, 2048) for r in pay])))


@pytest.mark.parametrize("n", [32, 64, 2048, 5632])
def test_dot_products(port, ref, n):
    rng = np.random.default_rng(100 + n)
    for trial in range(40):
        x = rand_rows(rng, 1, n, rng.choice([0.3, 1.0, 5.0]))[0]
        w = rand_rows(rng, 1, n, 0.02)
        xq, xh = ref.write_row(x, Q8), ref.write_row(x, F16)
        w8, w4, wh = W.quantize_payload(w, Q8), W.quantize_payload(w, Q4), W.quantize_payload(w, F16)
        for (a, adt, b, bdt) in [(xq, Q8, w8, Q8), (xq, Q8, w4, Q4), (xh, F16, wh, F16)]:
            assert port.vec_dot(a, adt, b, bdt, n).view(np.uint32) == ref.vec_dot(a, adt, b, bdt, n).view(np.uint32)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 63, 300, 2048])
def test_dot_f32(port, ref, n):
    rng = np.random.default_rng(n)
    for _ in range(20):
        a, b = rng.random(n).astype(np.float32), rand_rows(rng, 1, n)[0]
        assert port.vec_dot(a, F32, b, F32, n).view(np.uint32) == ref.vec_dot(a, F32, b, F32, n).view(np.uint32)


def test_libm_entry_points(port, ref):
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(-20, 20, 20000), rng.uniform(-104, 89, 5000), [0.0, -0.0, -np.inf, 88.7, -103.9]]).astype(np.float32)
    for x in xs:
        assert port.expf(x).view(np.uint32) == ref.expf(x).view(np.uint32)
    for pos in list(range(0, 64)) + [127, 128, 1000, 1535, 2047, 2175, 4095]:
        (c0, s0), (c1, s1) = port.rope_angles(pos, 64), ref.rope_angles(pos, 64)
        assert np.array_equal(bits(c0), bits(c1)) and np.array_equal(bits(s0), bits(s1)), pos


@pytest.mark.parametrize("wdt,adt", [(Q4, Q8), (Q8, Q8), (F16, F16)])
@pytest.mark.parametrize("shape", [(64, 2048), (96, 5632)])
def test_matmul(port, ref, wdt, adt, shape):
    n_out, k = shape
    rng = np.random.default_rng(n_out + wdt)
    n_ctx = 3
    x = ref.encode_rows(rand_rows(rng, n_ctx, k), adt)
    w = W.quantize_payload(rand_rows(rng, n_out, k, 0.02), wdt)
    for start in (0, 2):
        a = port.matmul_2d(x, adt, n_ctx, k, w, wdt, n_out, adt, start_pos=start)
        b = ref.matmul_2d(x, adt, n_ctx, k, w, wdt, n_out, adt, start_pos=start)
        assert np.array_equal(a[start:], b[start:])
    a = port.matmul_2d(x, adt, n_ctx, k, w, wdt, n_out, F32, out_1d=True, start_pos=n_ctx - 1)
    b = ref.matmul_2d(x, adt, n_ctx, k, w, wdt, n_out, F32, out_1d=True, start_pos=n_ctx - 1)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("adt", ADTS)
def test_elementwise_ops(port, ref, adt):
    rng = np.random.default_rng(7 + adt)
    n_ctx, n = 4, 2048
    x = ref.encode_rows(rand_rows(rng, n_ctx, n, 2.0), adt)
    y = ref.encode_rows(rand_rows(rng, n_ctx, n, 0.7), adt)
    wn = (1 + 0.1 * rng.standard_normal(n)).astype(np.float16)
    assert np.array_equal(port.rms_norm(x, adt, n_ctx, n, wn), ref.rms_norm(x, adt, n_ctx, n, wn))
    assert np.array_equal(port.silu(x, adt, n_ctx, n), ref.silu(x, adt, n_ctx, n))
    assert np.array_equal(port.mul(x, y, adt, n_ctx, n), ref.mul(x, y, adt, n_ctx, n))
    assert np.array_equal(port.add(x, y, adt, n_ctx, n), ref.add(x, y, adt, n_ctx, n))
    assert np.array_equal(port.rotary_emb(x, adt, n_ctx, n, 64), ref.rotary_emb(x, adt, n_ctx, n, 64))
    a, b = port.rotary_emb(x, adt, n_ctx, n, 64, start_pos=3), ref.rotary_emb(x, adt, n_ctx, n, 64, start_pos=3)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("wdt", [Q4, Q8, F16])
def test_token_embed(port, ref, wdt):
    rng = np.random.default_rng(9)
    n_vocab, n = 50, 2048
    adt = F16 if wdt == F16 else Q8
    w = W.quantize_payload(rand_rows(rng, n_vocab, n, 0.02), wdt)
    toks = rng.integers(0, n_vocab, 6).astype(np.int32)
    assert np.array_equal(port.token_embed(w, wdt, n_vocab, n, toks, adt), ref.token_embed(w, wdt, n_vocab, n, toks, adt))


@pytest.mark.parametrize("adt", ADTS)
@pytest.mark.parametrize("n_ctx,max_ctx", [(1, 64), (7, 64), (33, 128), (40, 128)])
def test_attention_inside_valid_domain(port, ref, adt, n_ctx, max_ctx):
    """Multi-row calls agree wherever the reference's score rows do not overlap."""
    assert (2 * n_ctx <= max_ctx) if adt == F16 else (((n_ctx + 31) // 32) * 34 <= max_ctx)
    rng = np.random.default_rng(n_ctx)
    H, G, D = 32, 4, 64
    q = ref.encode_rows(rand_rows(rng, n_ctx, H * D), adt)
    k = ref.encode_rows(rand_rows(rng, n_ctx, G * D), adt)
    v = ref.encode_rows(rand_rows(rng, n_ctx, G * D), adt)
    for start in {0, n_ctx - 1}:
        a = port.qkv_attn(q, k, v, adt, n_ctx, H, G, D, max_ctx, start_pos=start)
        b = ref.qkv_attn(q, k, v, adt, n_ctx, H, G, D, max_ctx, start_pos=start)
        assert np.array_equal(a[start:], b[start:])


def _mini_models(wdt, max_ctx, seed=3):
    cfg = W.mini_config(n_layers=2, n_vocab=256)
    wl = list(W.synth_weights(cfg, wdt, seed=seed))
    mp = oracle.port().model(cfg, max_ctx, wdt).load(wl)
    mr = oracle.ref().model(cfg, max_ctx, wdt).load(wl)
    return cfg, mp, mr


@pytest.mark.parametrize("wdt", [Q4, Q8, F16])
def test_mini_model_generate(ref, wdt):
    cfg, mp, mr = _mini_models(wdt, max_ctx=96)
    prompt = W.synth_prompt(11, 20, cfg.n_vocab)
    tp, _, lp = mp.generate(prompt, 12, want_logits=True)
    tr, _, lr = mr.generate(prompt, 12, want_logits=True)
    assert np.array_equal(tp, tr)
    assert np.array_equal(bits(lp), bits(lr))
    # every per-layer activation of the last row, raw encoded bytes
    row = prompt.size + 12 - 2
    for layer in range(cfg.n_layers):
        for name, aid in oracle.LAYER_ACVS.items():
            assert np.array_equal(mp.acv_raw(layer, aid, row), mr.acv_raw(layer, aid, row)), (layer, name)
    assert np.array_equal(mp.acv_raw(0, oracle.A_FINAL_NORM, row), mr.acv_raw(0, oracle.A_FINAL_NORM, row))
