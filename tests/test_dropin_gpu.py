"""Drop-in acceptance test: the UNMODIFIED reference application source (TinyLlama class, load_from_ckpt,
greedy_sample, print_perf) compiled against include/gten -- oracle/_ref/libdropin_refapp.so, built where /root/reference
exists -- loads a gten checkpoint and generates through the C++ module / op API on the GPU.  Its tokens and first logits
must equal the engine's (which the other tests pin bit-exactly to the reference CPU build)."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from tinyllama_cpp_b200 import weights as W

pytestmark = pytest.mark.gpu
LIB = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "libdropin_refapp.so"


@pytest.mark.skipif(not LIB.exists(), reason="oracle/_ref/libdropin_refapp.so not built (needs /root/reference at build time)")
def test_reference_app_over_dropin_headers(tmp_path):
    from tinyllama_cpp_b200 import capi
    capi.init(0)
    cfg = W.TINYLLAMA                                   # the reference class hard-codes these dimensions (tinyllama.cpp:12-20)
    wdt = W.Q4
    wl = list(W.synth_weights(cfg, wdt, seed=1))
    path = tmp_path / "tinyllama.q4.gten"
    W.write_gten(path, cfg, wdt, wl)
    prompt = W.synth_prompt(11, 9, cfg.n_vocab).astype(np.int32)
    n_new, max_ctx = 5, 32
    L = C.CDLL(str(LIB))
    L.dropin_generate.restype = C.c_int
    toks = np.zeros(n_new, np.int32)
    logits = np.zeros(cfg.n_vocab, np.float32)
    rc = L.dropin_generate(str(path).encode(), wdt, max_ctx, prompt.ctypes.data_as(C.c_void_p), prompt.size, n_new,
                           toks.ctypes.data_as(C.c_void_p), logits.ctypes.data_as(C.c_void_p))
    assert rc == 0
    e = capi.Engine(cfg, max_ctx, wdt).load(wl)
    want = e.generate(prompt, n_new)
    assert np.array_equal(toks, want[prompt.size:])
    assert np.array_equal(logits.view(np.uint32), e.logits(prompt, 0).view(np.uint32))
    e.close()


TOK = LIB.parent / "tokenizer.bin"


@pytest.mark.skipif(not (LIB.exists() and TOK.exists()), reason="oracle/_ref/libdropin_refapp.so or tokenizer.bin not built (needs /root/reference at build time)")
def test_reference_greedy_sample_and_print_perf(tmp_path):
    """The reference's OWN greedy_sample (tinyllama.cpp:395-440: Tokenizer.encode with the chat template, the per-token loop,
    Tokenizer.decode of every piece to stderr) and print_perf (:515-582), compiled unmodified over include/gten: the text it
    prints is the text of the engine's greedy tokens, and with gten::Timer in profiling mode the perf table shows device time."""
    from tinyllama_cpp_b200 import capi
    capi.init(0)
    cfg = W.TINYLLAMA
    wdt = W.Q8
    wl = list(W.synth_weights(cfg, wdt, seed=1))
    path = tmp_path / "tinyllama.q8.gten"
    W.write_gten(path, cfg, wdt, wl)
    L = C.CDLL(str(LIB))
    L.dropin_greedy_sample.restype = C.c_int
    L.dropin_decode.restype = C.c_int
    extra = 6
    text = C.create_string_buffer(1 << 16)
    ids = np.zeros(512, np.int32)
    n_ids = C.c_int()
    rc = L.dropin_greedy_sample(str(path).encode(), wdt, extra, str(TOK).encode(), b"Who is Karl Marx?", 1, text, len(text),
                                ids.ctypes.data_as(C.c_void_p), C.byref(n_ids))
    assert rc == 0
    n = n_ids.value
    # the first 15 ids of the golden comment (tinyllama.cpp:101-104) are the tokenizer's encoding of this prompt
    assert n >= 15
    out = text.value.decode(errors="replace")
    e = capi.Engine(cfg, n + extra, wdt).load(wl)
    toks = e.generate(ids[:n], extra)
    e.close()
    piece = C.create_string_buffer(1 << 14)
    L.dropin_decode(str(TOK).encode(), toks.astype(np.int32).ctypes.data_as(C.c_void_p), n, toks.size, piece, len(piece))
    want = piece.value.decode(errors="replace")
    assert out.startswith(want + "\n"), (out[:200], want)
    assert "PERFORMANCE" in out and "Lin time [per tok]" in out
    total = [ln for ln in out.splitlines() if "Inference [total]" in ln]
    assert total and int(total[0].split(":")[1].replace("ms", "").strip()) > 0, total     # device time, not launch overhead
