"""Pins tinyllama_cpp_b200.weights (the vectorised restatement of the reference converter) against the reference's OWN code:
the function definitions of /root/reference/tinyllama_to_gten.py:1-148 (itob, q8_quantize, q4_quantize, write_layer) are
exec'd unmodified -- everything before its argparse block -- and fed the same tensors; payloads and whole records must be
byte-identical.  Runs only where /root/reference exists (this container); the converter is the normative spec of the
on-disk format and of WEIGHT quantisation (ties-to-even torch.round, unlike the activation quantiser's roundf)."""
import io
import struct
from pathlib import Path

import numpy as np
import pytest

from tinyllama_cpp_b200 import weights as W

REF = Path("/root/reference/tinyllama_to_gten.py")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="/root/reference absent (GPU box): the pin runs in the build container")


@pytest.fixture(scope="module")
def ref_ns():
    torch = pytest.importorskip("torch")
    src = REF.read_text()
    head = src[: src.index("parser = argparse.ArgumentParser")] if "parser = argparse.ArgumentParser" in src else src[: src.index("def convert")]
    ns = {}
    exec(compile(head, str(REF), "exec"), ns)
    ns["torch"] = torch
    return ns


def _cases():
    rng = np.random.default_rng(77)
    a = (rng.standard_normal((24, 256)) * 0.02).astype(np.float32)
    b = a.copy()
    b[3, 32:64] = 0.0                                   # an all-zero block: delta 0, Q4 nibbles 0x77
    b[5, :32] = np.linspace(-1, 1, 32, dtype=np.float32) * np.float32(0.5)
    # exact .5 ties in x * (1 / delta): absmax 127 -> delta 1, scale 1; halves must round to EVEN (torch.round)
    c = np.zeros((2, 64), np.float32)
    c[0, :32] = np.array([127.0, 0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 3.5] * 4, np.float32)
    c[1, :32] = np.array([7.0, 0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 3.5] * 4, np.float32)
    d = W.synth_tensor(1, 0, W.T_Q, 2048, 2048, 0, 16)  # the generator the golden runs use
    return {"gauss": a, "zero_block": b, "ties": c, "synthetic": d}


@pytest.mark.parametrize("name", list(_cases()))
def test_quantisers_match_reference_converter(ref_ns, name):
    torch = ref_ns["torch"]
    w = _cases()[name]
    t = torch.from_numpy(w.copy())
    d8, q8 = ref_ns["q8_quantize"](t.clone())
    md8, mq8 = W.q8_quantize(w)
    assert np.array_equal(d8.numpy().view(np.uint16), md8.view(np.uint16).reshape(-1))
    assert np.array_equal(q8.numpy(), mq8)
    d4, q4 = ref_ns["q4_quantize"](t.clone())
    md4, mq4 = W.q4_quantize(w)
    assert np.array_equal(d4.numpy().view(np.uint16), md4.view(np.uint16).reshape(-1))
    assert np.array_equal(q4.numpy(), mq4)


@pytest.mark.parametrize("dtype,wdt", [("fp16", W.F16), ("q8", W.Q8), ("q4", W.Q4)])
def test_record_bytes_match_reference_write_layer(ref_ns, tmp_path, dtype, wdt):
    """One record written by the reference's write_layer == the same record from write_gten (header, name twice, size, payload)."""
    torch = ref_ns["torch"]
    cfg = W.mini_config(n_layers=1, n_vocab=64)
    w = _cases()["gauss"]
    name = "model.layers.0.self_attn.q_proj.weight"
    buf = io.BytesIO()
    ref_ns["write_layer"](buf, name, torch.from_numpy(w.copy()), dtype)
    want = buf.getvalue()
    payload = W.quantize_payload(w, wdt)
    enc = name.encode()
    mine = struct.pack("<i", len(enc)) + enc + struct.pack("<i", len(enc)) + enc + struct.pack("<i", payload.size) + payload.tobytes()
    assert mine == want
    # and through write_gten / read_gten: magic, then the records in the converter's tensor order (tinyllama_to_gten.py:157-201)
    wl = list(W.synth_weights(cfg, wdt, seed=5))
    path = tmp_path / f"m.{dtype}.gten"
    W.write_gten(path, cfg, wdt, wl)
    raw = path.read_bytes()
    assert struct.unpack("<q", raw[:8])[0] == ref_ns["GTEN_MAGIC_NUMBER"]
    first = W.tensor_list(cfg)[0][2].encode()
    assert raw[8:12] == ref_ns["itob"](len(first)) and raw[12:12 + len(first)] == first
    back = list(W.read_gten(path, cfg, wdt))
    assert all(np.array_equal(a[2], np.ascontiguousarray(b[2]).view(np.uint8).reshape(-1)) for a, b in zip(back, wl))
