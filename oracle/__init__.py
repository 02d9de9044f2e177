"""oracle/ -- CPU checkers for the tinyllama.cpp forward hot path.  TEST INFRASTRUCTURE ONLY.

Two libraries with the same C interface (prefix aside):

* ``port()``  -> oracle/liboracle.so, a plain-C restatement (oracle/gten_oracle.c, prefix ``orc_``);
* ``ref()``   -> oracle/_ref/libgten_ref.so, the UNMODIFIED reference compiled from /root/reference
  behind oracle/ref_harness.cpp (prefix ``ref_``).  Built only where /root/reference exists; the
  prebuilt file travels to the GPU box.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package.  The product (tinyllama.cpp_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent

I32, F16, F32, Q8, Q4 = 0, 1, 2, 3, 4           # gten_types.h:20-26
DTYPE_NAMES = {I32: "i32", F16: "f16", F32: "f32", Q8: "q8", Q4: "q4"}

# tensor ids / activation ids (ref_harness.cpp, gten_oracle.h, include/gten_b200.h)
T_EMBED, T_FINAL_NORM, T_LM_HEAD = 0, 1, 2
T_Q, T_K, T_V, T_O, T_GATE, T_UP, T_DOWN, T_ATTN_NORM, T_FFN_NORM = 10, 11, 12, 13, 14, 15, 16, 17, 18
A_EMB, A_FINAL_NORM = 0, 1
(A_ATTN_NORM, A_Q, A_K, A_V, A_ATTN_OUT, A_O, A_INP_RES, A_FFN_NORM, A_GATE, A_UP, A_DOWN, A_ATTN_RES) = range(10, 22)
LAYER_ACVS = {
    "attn_norm": A_ATTN_NORM, "q": A_Q, "k": A_K, "v": A_V, "attn_out": A_ATTN_OUT, "o": A_O,
    "inp_res": A_INP_RES, "ffn_norm": A_FFN_NORM, "gate": A_GATE, "up": A_UP, "down": A_DOWN,
    "attn_res": A_ATTN_RES,
}


def row_nbytes(dtype: int, n: int) -> int:
    """Byte size of one row of n elements (tensor.cpp:37-58)."""
    if dtype == Q8:
        return ((n + 31) // 32) * 34
    if dtype == Q4:
        assert n % 32 == 0
        return (n // 32) * 18
    if dtype == F16:
        return n * 2
    return n * 4


def build(verbose: bool = False) -> None:
    """Compile liboracle.so and, when /root/reference is present, _ref/libgten_ref.so."""
    r = subprocess.run(["make", "-C", str(HERE), "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)
    return a


class CpuLib:
    """ctypes view of one of the two CPU checkers; every method takes/returns numpy arrays."""

    def __init__(self, path: Path, prefix: str, kind: str):
        self.path, self.prefix, self.kind = path, prefix, kind
        self.lib = C.CDLL(str(path))
        L, p = self.lib, prefix
        vp, i, f = C.c_void_p, C.c_int, C.c_float

        def sig(name, res, *args):
            fn = getattr(L, p + name)
            fn.restype, fn.argtypes = res, list(args)
            setattr(self, "_" + name, fn)

        sig("build_info", C.c_char_p)
        sig("fp32_to_fp16", C.c_uint16, f)
        sig("fp16_to_fp32", f, C.c_uint16)
        sig("q8_quantize_row", None, vp, vp, i)
        sig("q8_dequantize_row", None, vp, vp, i)
        sig("q4_dequantize_row", None, vp, vp, i)
        sig("read_row_to_float", None, vp, i, vp, i)
        sig("write_row_from_float", None, vp, vp, i, i)
        sig("vec_dot_product", f, vp, i, vp, i, i)
        sig("token_embed", None, vp, i, i, i, vp, i, vp, i, i)
        sig("matmul_2d", None, vp, i, i, i, vp, i, i, vp, i, i, i)
        sig("rms_norm", None, vp, i, i, i, vp, vp, i)
        sig("rotary_emb", None, vp, i, i, i, i, i)
        sig("silu", None, vp, i, i, i, vp, i)
        sig("mul", None, vp, vp, i, i, i, vp, i)
        sig("add", None, vp, vp, i, i, i, vp, i)
        sig("qkv_attn", None, vp, vp, vp, vp, vp, i, i, i, i, i, i, i)
        sig("expf", f, f)
        if prefix == "orc_":
            sig("expf_bits_range", None, C.c_uint32, C.c_uint32, vp)
        sig("rope_angles", None, i, i, vp, vp)
        sig("model_new", vp, i, i, i, i, i, i, i, i)
        sig("model_free", None, vp)
        sig("model_weight", vp, vp, i, i, C.POINTER(C.c_int64))
        sig("model_logits", None, vp, vp, i, i, vp)
        sig("model_generate", None, vp, vp, i, i, vp, vp)
        sig("model_acv", i, vp, i, i, i, vp)
        sig("model_acv_raw", i, vp, i, i, i, vp)
        if prefix == "orc_":
            sig("model_capture_row", None, vp, i)

    def build_info(self) -> str:
        return self._build_info().decode()

    # ---- rows -------------------------------------------------------------------------
    def fp32_to_fp16(self, x) -> int:
        return int(self._fp32_to_fp16(float(np.float32(x))))

    def fp16_to_fp32(self, h) -> np.float32:
        return np.float32(self._fp16_to_fp32(int(h)))

    def write_row(self, x: np.ndarray, dtype: int) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32).copy()
        out = np.zeros(row_nbytes(dtype, x.size), dtype=np.uint8)
        self._write_row_from_float(_ptr(x), _ptr(out), dtype, x.size)
        return out

    def read_row(self, raw: np.ndarray, dtype: int, n: int) -> np.ndarray:
        raw = np.ascontiguousarray(raw)
        out = np.zeros(n, dtype=np.float32)
        self._read_row_to_float(_ptr(raw), dtype, _ptr(out), n)
        return out

    def encode_rows(self, x: np.ndarray, dtype: int) -> np.ndarray:
        """[rows, n] fp32 -> [rows, row_nbytes] uint8."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        return np.stack([self.write_row(r, dtype) for r in x])

    def decode_rows(self, raw: np.ndarray, dtype: int, n: int) -> np.ndarray:
        return np.stack([self.read_row(r, dtype, n) for r in raw])

    def vec_dot(self, a, adt, b, bdt, n) -> np.float32:
        return np.float32(self._vec_dot_product(_ptr(np.ascontiguousarray(a)), adt, _ptr(np.ascontiguousarray(b)), bdt, n))

    def expf(self, x) -> np.float32:
        return np.float32(self._expf(float(np.float32(x))))

    def expf_bits_range(self, first: int, count: int) -> np.ndarray:
        """Host libm expf of the float bit patterns [first, first + count) (port library only: it is libm, not a restatement)."""
        out = np.empty(count, np.float32)
        self._expf_bits_range(first, count, _ptr(out))
        return out

    def rope_angles(self, pos: int, d_head: int):
        c = np.zeros(d_head // 2, np.float32)
        s = np.zeros(d_head // 2, np.float32)
        self._rope_angles(pos, d_head, _ptr(c), _ptr(s))
        return c, s

    # ---- ops (raw uint8 row-major buffers in, new buffer out) --------------------------
    def token_embed(self, w, wdt, n_vocab, n_embd, tokens, odt, start_pos=0):
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.zeros((tokens.size, row_nbytes(odt, n_embd)), np.uint8)
        self._token_embed(_ptr(w), wdt, n_vocab, n_embd, _ptr(tokens), tokens.size, _ptr(out), odt, start_pos)
        return out

    def matmul_2d(self, x, xdt, n_ctx, k, w, wdt, n_out, odt, out_1d=False, start_pos=0):
        out = np.zeros(row_nbytes(odt, n_out) if out_1d else (n_ctx, row_nbytes(odt, n_out)), np.uint8)
        self._matmul_2d(_ptr(x), xdt, n_ctx, k, _ptr(w), wdt, n_out, _ptr(out), odt, int(out_1d), start_pos)
        return out

    def rms_norm(self, x, xdt, n_ctx, n_embd, w_f16, start_pos=0):
        out = np.zeros_like(x)
        self._rms_norm(_ptr(x), xdt, n_ctx, n_embd, _ptr(np.ascontiguousarray(w_f16)), _ptr(out), start_pos)
        return out

    def rotary_emb(self, x, xdt, n_ctx, n_embd, d_head, start_pos=0):
        out = np.ascontiguousarray(x).copy()
        self._rotary_emb(_ptr(out), xdt, n_ctx, n_embd, d_head, start_pos)
        return out

    def silu(self, x, xdt, n_ctx, n_embd, start_pos=0):
        out = np.zeros_like(x)
        self._silu(_ptr(x), xdt, n_ctx, n_embd, _ptr(out), start_pos)
        return out

    def mul(self, a, b, xdt, n_ctx, n_embd, start_pos=0):
        out = np.zeros_like(a)
        self._mul(_ptr(a), _ptr(b), xdt, n_ctx, n_embd, _ptr(out), start_pos)
        return out

    def add(self, a, b, xdt, n_ctx, n_embd, start_pos=0):
        out = np.zeros_like(a)
        self._add(_ptr(a), _ptr(b), xdt, n_ctx, n_embd, _ptr(out), start_pos)
        return out

    def qkv_attn(self, q, k, v, xdt, n_ctx, n_heads, n_kv, d_head, max_ctx, start_pos=0):
        out = np.zeros_like(q)
        # score scratch sized like SelfAttention's qk_acv (modules.cpp:180)
        qk = np.zeros(n_heads * max_ctx * row_nbytes(xdt, max_ctx) + 64, np.uint8)
        self._qkv_attn(_ptr(q), _ptr(k), _ptr(v), _ptr(qk), _ptr(out), xdt, n_ctx, n_heads, n_kv, d_head, max_ctx, start_pos)
        return out

    # ---- model -------------------------------------------------------------------------
    def model(self, cfg, max_ctx: int, wdtype: int) -> "CpuModel":
        return CpuModel(self, cfg, max_ctx, wdtype)


class CpuModel:
    """A TinyLlama-shaped network on one of the CPU checkers (weights are injected as gten payloads)."""

    def __init__(self, lib: CpuLib, cfg, max_ctx: int, wdtype: int):
        self.lib, self.cfg, self.max_ctx, self.wdtype = lib, cfg, max_ctx, wdtype
        self.adtype = F16 if wdtype == F16 else Q8
        self.h = lib._model_new(cfg.n_vocab, cfg.n_embd, cfg.n_ffn, cfg.n_layers, cfg.n_heads, cfg.n_groups, max_ctx, wdtype)
        assert self.h

    def close(self):
        if self.h:
            self.lib._model_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weight(self, layer: int, tid: int, payload: np.ndarray):
        nb = C.c_int64(0)
        p = self.lib._model_weight(self.h, layer, tid, C.byref(nb))
        payload = np.ascontiguousarray(payload).view(np.uint8).reshape(-1)
        assert p and nb.value == payload.size, f"weight {layer}/{tid}: oracle expects {nb.value} B, got {payload.size}"
        C.memmove(p, payload.ctypes.data, payload.size)

    def load(self, weights) -> "CpuModel":
        """weights: iterable of (layer, tensor_id, payload uint8 array)."""
        for layer, tid, payload in weights:
            self.set_weight(layer, tid, payload)
        return self

    def logits(self, tokens, start_pos: int) -> np.ndarray:
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.zeros(self.cfg.n_vocab, np.float32)
        self.lib._model_logits(self.h, _ptr(tokens), tokens.size, start_pos, _ptr(out))
        return out

    def generate(self, prompt, n_new: int, want_logits: bool = False):
        """Greedy loop of greedy_sample (tinyllama.cpp:395-440), no EOS stop.
        Returns (tokens[n_prompt+n_new], (prefill_s, decode_s), logits or None)."""
        prompt = np.asarray(prompt, dtype=np.int32)
        toks = np.zeros(prompt.size + n_new, np.int32)
        toks[: prompt.size] = prompt
        times = np.zeros(2, np.float64)
        lg = np.zeros((n_new, self.cfg.n_vocab), np.float32) if want_logits else None
        self.lib._model_generate(self.h, _ptr(toks), prompt.size, n_new, _ptr(times), _ptr(lg))
        return toks, (float(times[0]), float(times[1])), lg

    def capture_row(self, row: int):
        if self.lib.prefix == "orc_":
            self.lib._model_capture_row(self.h, row)

    def acv(self, layer: int, aid: int, row: int) -> np.ndarray:
        buf = np.zeros(max(self.cfg.n_ffn, self.cfg.n_embd), np.float32)
        w = self.lib._model_acv(self.h, layer, aid, row, _ptr(buf))
        assert w > 0, f"acv({layer},{aid},{row}) -> {w}"
        return buf[:w].copy()

    def acv_raw(self, layer: int, aid: int, row: int) -> np.ndarray:
        buf = np.zeros(4 * max(self.cfg.n_ffn, self.cfg.n_embd), np.uint8)
        nb = self.lib._model_acv_raw(self.h, layer, aid, row, _ptr(buf))
        assert nb > 0, f"acv_raw({layer},{aid},{row}) -> {nb}"
        return buf[:nb].copy()


_cache: dict = {}


def port() -> CpuLib:
    if "port" not in _cache:
        so = HERE / "liboracle.so"
        if not so.exists():
            build()
        _cache["port"] = CpuLib(so, "orc_", "port")
    return _cache["port"]


def ref_available() -> bool:
    return (HERE / "_ref" / "libgten_ref.so").exists()


def ref() -> CpuLib:
    if "ref" not in _cache:
        so = HERE / "_ref" / "libgten_ref.so"
        if not so.exists() and Path("/root/reference/tinyllama.cpp").exists():
            build()
        if not so.exists():
            raise FileNotFoundError("oracle/_ref/libgten_ref.so is missing and /root/reference is absent")
        _cache["ref"] = CpuLib(so, "ref_", "reference")
    return _cache["ref"]


def best() -> CpuLib:
    """The real reference when its prebuilt library is present, else the C restatement."""
    return ref() if ref_available() else port()
