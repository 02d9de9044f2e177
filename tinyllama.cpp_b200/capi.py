"""ctypes binding of libgten_b200.so (include/gten_b200.h) -- used by tests/, bench.py and smoke().

No compute lives here and there is no fallback: if the shared library is missing, or no CUDA device is
present, every call raises.  numpy arrays are host buffers; DeviceBuffer wraps gtb_malloc storage.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import weights as W

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libgten_b200.so"

I32, F16, F32, Q8, Q4 = 0, 1, 2, 3, 4

# every symbol include/gten_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "gtb_last_error", "gtb_version", "gtb_init", "gtb_device_count", "gtb_device_info", "gtb_sync", "gtb_stream",
    "gtb_launch_count", "gtb_mem_allocated", "gtb_malloc", "gtb_free", "gtb_memset", "gtb_h2d", "gtb_d2h", "gtb_d2d",
    "gtb_host_alloc", "gtb_host_free", "gtb_weight_upload", "gtb_weight_from_device", "gtb_weight_free",
    "gtb_weight_nbytes", "gtb_weight_dequant", "gtb_write_rows_from_float", "gtb_read_rows_to_float",
    "gtb_token_embed", "gtb_matmul_2d", "gtb_rms_norm", "gtb_rotary_emb", "gtb_silu", "gtb_mul", "gtb_add",
    "gtb_qkv_attn", "gtb_engine_create", "gtb_engine_destroy", "gtb_engine_set_weight", "gtb_engine_load_gten",
    "gtb_engine_logits", "gtb_engine_generate", "gtb_engine_reset", "gtb_engine_prefill", "gtb_engine_decode",
    "gtb_engine_position", "gtb_engine_read_tokens", "gtb_engine_read_logits", "gtb_engine_acv",
    "gtb_engine_set_option", "gtb_engine_weight_bytes", "gtb_engine_read_prof", "gtb_selftest_exact_sum", "gtb_engine_uses_megakernel",
    "gtb_engine_prefill_fast", "gtb_engine_pf_acv", "gtb_pf_gemm_f32", "gtb_engine_topk",
    "gtb_engine_batch_create", "gtb_engine_batch_adopt", "gtb_engine_batch_decode", "gtb_engine_batch_position",
    "gtb_engine_batch_read_tokens", "gtb_engine_batch_read_logits", "gtb_engine_batch_prefill", "gtb_selftest_expf", "gtb_vec_dot_product",
]


class GtbError(RuntimeError):
    pass


class ModelConfigC(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("n_vocab", "n_embd", "n_ffn", "n_layers", "n_heads", "n_groups", "max_ctx", "wdtype")]


_lib = None


def lib():
    """The loaded shared library (raises if it has not been built: there is no Python/CPU fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise GtbError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(str(LIB_PATH))
        vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
        L.gtb_last_error.restype = C.c_char_p
        L.gtb_version.restype = C.c_char_p
        L.gtb_stream.restype = vp
        L.gtb_launch_count.restype = C.c_int64
        L.gtb_mem_allocated.restype = C.c_int64
        sigs = {
            "gtb_init": [i], "gtb_device_count": [C.POINTER(i)], "gtb_device_info": [C.POINTER(i)] * 3 + [C.POINTER(sz)],
            "gtb_sync": [], "gtb_malloc": [C.POINTER(vp), sz], "gtb_free": [vp], "gtb_memset": [vp, i, sz],
            "gtb_h2d": [vp, vp, sz], "gtb_d2h": [vp, vp, sz], "gtb_d2d": [vp, vp, sz],
            "gtb_host_alloc": [C.POINTER(vp), sz], "gtb_host_free": [vp],
            "gtb_weight_upload": [C.POINTER(vp), vp, i, i, i], "gtb_weight_from_device": [C.POINTER(vp), vp, i, i, i],
            "gtb_weight_free": [vp], "gtb_weight_nbytes": [vp, C.POINTER(sz)], "gtb_weight_dequant": [vp, i, i, vp],
            "gtb_write_rows_from_float": [vp, vp, i, i, i], "gtb_read_rows_to_float": [vp, i, vp, i, i],
            "gtb_token_embed": [vp, vp, vp, i, i, i], "gtb_matmul_2d": [vp, i, i, vp, vp, i, i, i],
            "gtb_rms_norm": [vp, i, i, i, vp, vp, i], "gtb_rotary_emb": [vp, i, i, i, i, i],
            "gtb_silu": [vp, i, i, i, vp, i], "gtb_mul": [vp, vp, i, i, i, vp, i], "gtb_add": [vp, vp, i, i, i, vp, i],
            "gtb_qkv_attn": [vp, vp, vp, vp, vp, i, i, i, i, i, i, i],
            "gtb_engine_create": [C.POINTER(vp), C.POINTER(ModelConfigC)], "gtb_engine_destroy": [vp],
            "gtb_engine_set_weight": [vp, i, i, vp, sz], "gtb_engine_load_gten": [vp, C.c_char_p],
            "gtb_engine_logits": [vp, vp, i, i, vp], "gtb_engine_generate": [vp, vp, i, i, i, C.POINTER(i)],
            "gtb_engine_reset": [vp], "gtb_engine_prefill": [vp, vp, i], "gtb_engine_decode": [vp, i],
            "gtb_engine_position": [vp, C.POINTER(i)], "gtb_engine_read_tokens": [vp, vp, i, i],
            "gtb_engine_read_logits": [vp, vp], "gtb_engine_acv": [vp, i, i, vp, C.POINTER(i)],
            "gtb_engine_set_option": [vp, C.c_char_p, i], "gtb_engine_weight_bytes": [vp, C.POINTER(sz)],
            "gtb_engine_read_prof": [vp, vp, i], "gtb_selftest_exact_sum": [vp, i, vp], "gtb_engine_uses_megakernel": [vp, C.POINTER(C.c_int)],
            "gtb_engine_prefill_fast": [vp, vp, i], "gtb_engine_pf_acv": [vp, i, i, i, vp, C.POINTER(i)],
            "gtb_pf_gemm_f32": [vp, vp, i, i, i, i, vp], "gtb_engine_topk": [vp, i, vp, vp],
            "gtb_engine_batch_create": [vp, i], "gtb_engine_batch_adopt": [vp, i], "gtb_engine_batch_decode": [vp, i],
            "gtb_engine_batch_position": [vp, i, C.POINTER(i)], "gtb_engine_batch_read_tokens": [vp, i, vp, i, i],
            "gtb_engine_batch_read_logits": [vp, i, vp], "gtb_engine_batch_prefill": [vp, i, vp, i],
            "gtb_selftest_expf": [C.c_uint32, C.c_uint32, vp],
            "gtb_vec_dot_product": [vp, i, vp, i, i, C.POINTER(C.c_float)],
        }
        for name, args in sigs.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise GtbError(f"gtb error {rc}: {lib().gtb_last_error().decode()}")


def init(device: int = 0):
    check(lib().gtb_init(device))


def sync():
    check(lib().gtb_sync())


def launch_count() -> int:
    return int(lib().gtb_launch_count())


def device_info():
    sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
    check(lib().gtb_device_info(C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
    return {"sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value}


def stream_handle() -> int:
    return int(lib().gtb_stream() or 0)


def _hp(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def pf_gemm_f32(a16: np.ndarray, w16: np.ndarray, bn: int = 256) -> np.ndarray:
    """C = A . W^T on the tcgen05 path (self-test entry): A [M,K] fp16, W [N,K] fp16 -> C [M,N] fp32."""
    a16 = np.ascontiguousarray(a16, np.float16)
    w16 = np.ascontiguousarray(w16, np.float16)
    m, k = a16.shape
    n = w16.shape[0]
    out = np.empty((m, n), np.float32)
    check(lib().gtb_pf_gemm_f32(_hp(a16), _hp(w16), m, n, k, bn, _hp(out)))
    return out


class DeviceBuffer:
    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        self.nbytes = int(nbytes)
        check(lib().gtb_malloc(C.byref(self.ptr), self.nbytes))

    @classmethod
    def from_host(cls, a: np.ndarray) -> "DeviceBuffer":
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        check(lib().gtb_h2d(b.ptr, _hp(a), a.nbytes))
        sync()      # `a` may be a temporary
        return b

    def to_host(self, dtype=np.uint8, shape=None) -> np.ndarray:
        out = np.empty(self.nbytes, np.uint8)
        check(lib().gtb_d2h(_hp(out), self.ptr, self.nbytes))
        out = out.view(dtype)
        return out.reshape(shape) if shape is not None else out

    def free(self):
        if self.ptr:
            lib().gtb_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Weight:
    """A weight matrix in the device layout (one-time repack of the gten payload)."""

    def __init__(self, payload: np.ndarray, dtype: int, rows: int, cols: int):
        self.h = C.c_void_p()
        self.dtype, self.rows, self.cols = dtype, rows, cols
        payload = np.ascontiguousarray(payload).view(np.uint8).reshape(-1)
        assert payload.size == rows * W.row_nbytes(dtype, cols)
        check(lib().gtb_weight_upload(C.byref(self.h), _hp(payload), dtype, rows, cols))

    def dequant(self, row0: int = 0, nrows: int | None = None) -> np.ndarray:
        nrows = self.rows - row0 if nrows is None else nrows
        out = np.empty((nrows, self.cols), np.float32)
        check(lib().gtb_weight_dequant(self.h, row0, nrows, _hp(out)))
        return out

    def nbytes(self) -> int:
        n = C.c_size_t()
        check(lib().gtb_weight_nbytes(self.h, C.byref(n)))
        return n.value

    def free(self):
        if self.h:
            lib().gtb_weight_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- op-level helpers: numpy in (reference row layout), numpy out -----------------------------------
def write_rows(x: np.ndarray, dtype: int) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    rows, n = x.shape
    din, dout = DeviceBuffer.from_host(x), DeviceBuffer(rows * W.row_nbytes(dtype, n))
    check(lib().gtb_memset(dout.ptr, 0, dout.nbytes))     # bytes past a partial last block are never written
    check(lib().gtb_write_rows_from_float(din.ptr, dout.ptr, dtype, rows, n))
    return dout.to_host(np.uint8, (rows, -1))


def read_rows(raw: np.ndarray, dtype: int, n: int) -> np.ndarray:
    raw = np.ascontiguousarray(raw)
    rows = raw.shape[0]
    din, dout = DeviceBuffer.from_host(raw), DeviceBuffer(rows * n * 4)
    check(lib().gtb_read_rows_to_float(din.ptr, dtype, dout.ptr, rows, n))
    return dout.to_host(np.float32, (rows, n))


def token_embed(w: Weight, tokens: np.ndarray, odt: int, start_pos: int = 0) -> np.ndarray:
    tokens = np.ascontiguousarray(tokens, np.int32)
    dt, dout = DeviceBuffer.from_host(tokens), DeviceBuffer(tokens.size * W.row_nbytes(odt, w.cols))
    check(lib().gtb_memset(dout.ptr, 0, dout.nbytes))
    check(lib().gtb_token_embed(w.h, dt.ptr, dout.ptr, odt, tokens.size, start_pos))
    return dout.to_host(np.uint8, (tokens.size, -1))


def matmul_2d(x: np.ndarray, xdt: int, n_ctx: int, w: Weight, odt: int, out_1d: bool = False, start_pos: int = 0) -> np.ndarray:
    dx = DeviceBuffer.from_host(x)
    rows_out = 1 if out_1d else n_ctx
    dout = DeviceBuffer(rows_out * W.row_nbytes(odt, w.rows))
    check(lib().gtb_memset(dout.ptr, 0, dout.nbytes))
    check(lib().gtb_matmul_2d(dx.ptr, xdt, n_ctx, w.h, dout.ptr, odt, int(out_1d), start_pos))
    out = dout.to_host(np.uint8)
    return out if out_1d else out.reshape(n_ctx, -1)


def _rowwise(fn, x, dtype, n_ctx, n, start_pos, second=None, extra=()):
    dx = DeviceBuffer.from_host(x)
    dout = DeviceBuffer(dx.nbytes)
    check(lib().gtb_memset(dout.ptr, 0, dout.nbytes))
    if second is not None:
        db = DeviceBuffer.from_host(second)
        check(fn(dx.ptr, db.ptr, dtype, n_ctx, n, dout.ptr, start_pos))
    else:
        check(fn(dx.ptr, dtype, n_ctx, n, *extra, dout.ptr, start_pos))
    return dout.to_host(np.uint8, x.shape)


def rms_norm(x, dtype, n_ctx, n, w_f16, start_pos=0):
    dw = DeviceBuffer.from_host(np.ascontiguousarray(w_f16))
    return _rowwise(lib().gtb_rms_norm, x, dtype, n_ctx, n, start_pos, extra=(dw.ptr,))


def silu(x, dtype, n_ctx, n, start_pos=0):
    return _rowwise(lib().gtb_silu, x, dtype, n_ctx, n, start_pos)


def mul(a, b, dtype, n_ctx, n, start_pos=0):
    return _rowwise(lib().gtb_mul, a, dtype, n_ctx, n, start_pos, second=b)


def add(a, b, dtype, n_ctx, n, start_pos=0):
    return _rowwise(lib().gtb_add, a, dtype, n_ctx, n, start_pos, second=b)


def rotary_emb(x, dtype, n_ctx, n, d_head, start_pos=0):
    dx = DeviceBuffer.from_host(x)
    check(lib().gtb_rotary_emb(dx.ptr, dtype, n_ctx, n, d_head, start_pos))
    return dx.to_host(np.uint8, x.shape)


def qkv_attn(q, k, v, dtype, n_ctx, n_heads, n_kv, d_head, max_ctx, start_pos=0):
    dq, dk, dv = DeviceBuffer.from_host(q), DeviceBuffer.from_host(k), DeviceBuffer.from_host(v)
    dout = DeviceBuffer(dq.nbytes)
    check(lib().gtb_memset(dout.ptr, 0, dout.nbytes))
    check(lib().gtb_qkv_attn(dq.ptr, dk.ptr, dv.ptr, None, dout.ptr, dtype, n_ctx, n_heads, n_kv, d_head, max_ctx, start_pos))
    return dout.to_host(np.uint8, q.shape)


# ---- engine -----------------------------------------------------------------------------------------
def selftest_exact_sum(terms, cycles: bool = False):
    terms = np.ascontiguousarray(terms, np.float32)
    out = np.zeros(8, np.float32)
    check(lib().gtb_selftest_exact_sum(_hp(terms), terms.size, _hp(out)))
    if cycles:
        return out[0], out[1:5].copy()
    return out[0]


def vec_dot_product(a: np.ndarray, adt: int, b: np.ndarray, bdt: int, n: int) -> np.float32:
    """ops::vec_dot_product on two host rows in the reference layout (gten/ops.h:482-512)."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    out = C.c_float()
    check(lib().gtb_vec_dot_product(_hp(a), adt, _hp(b), bdt, n, C.byref(out)))
    return np.float32(out.value)


def selftest_expf(first_bits: int, count: int) -> np.ndarray:
    """Device expf of the float bit patterns [first_bits, first_bits + count)."""
    out = np.empty(count, np.float32)
    check(lib().gtb_selftest_expf(first_bits, count, _hp(out)))
    return out


class Engine:
    """TinyLlama{max_ctx, dtype} resident on the GPU (tinyllama.cpp:23-76)."""

    def __init__(self, cfg: W.ModelConfig, max_ctx: int, wdtype: int):
        self.cfg, self.max_ctx, self.wdtype = cfg, max_ctx, wdtype
        self.h = C.c_void_p()
        c = ModelConfigC(cfg.n_vocab, cfg.n_embd, cfg.n_ffn, cfg.n_layers, cfg.n_heads, cfg.n_groups, max_ctx, wdtype)
        check(lib().gtb_engine_create(C.byref(self.h), C.byref(c)))

    def close(self):
        if self.h:
            lib().gtb_engine_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weight(self, layer: int, tid: int, payload: np.ndarray):
        payload = np.ascontiguousarray(payload).view(np.uint8).reshape(-1)
        check(lib().gtb_engine_set_weight(self.h, layer, tid, _hp(payload), payload.size))

    def load(self, weights) -> "Engine":
        for layer, tid, payload in weights:
            self.set_weight(layer, tid, payload)
        return self

    def load_gten(self, path) -> "Engine":
        check(lib().gtb_engine_load_gten(self.h, str(path).encode()))
        return self

    def set_option(self, name: str, value: int):
        check(lib().gtb_engine_set_option(self.h, name.encode(), int(value)))

    def logits(self, tokens, start_pos: int) -> np.ndarray:
        tokens = np.ascontiguousarray(tokens, np.int32)
        out = np.empty(self.cfg.n_vocab, np.float32)
        check(lib().gtb_engine_logits(self.h, _hp(tokens), tokens.size, start_pos, _hp(out)))
        return out

    def generate(self, prompt, n_new: int, eos_id: int = -1) -> np.ndarray:
        prompt = np.asarray(prompt, np.int32)
        toks = np.zeros(prompt.size + n_new, np.int32)
        toks[: prompt.size] = prompt
        n = C.c_int()
        check(lib().gtb_engine_generate(self.h, _hp(toks), prompt.size, n_new, eos_id, C.byref(n)))
        return toks[: prompt.size + n.value]

    def prefill(self, tokens):
        tokens = np.ascontiguousarray(tokens, np.int32)
        check(lib().gtb_engine_prefill(self.h, _hp(tokens), tokens.size))

    def prefill_fast(self, tokens):
        """Batched prefill (tcgen05 GEMMs); tolerance-level parity, K/V cache and first token left for decode()."""
        tokens = np.ascontiguousarray(tokens, np.int32)
        check(lib().gtb_engine_prefill_fast(self.h, _hp(tokens), tokens.size))

    def pf_acv(self, layer: int, aid: int, row: int) -> np.ndarray:
        out = np.empty(max(self.cfg.n_ffn, self.cfg.n_embd), np.float32)
        w = C.c_int()
        check(lib().gtb_engine_pf_acv(self.h, layer, aid, row, _hp(out), C.byref(w)))
        return out[: w.value].copy()

    def decode(self, n_steps: int):
        check(lib().gtb_engine_decode(self.h, n_steps))

    def position(self) -> int:
        p = C.c_int()
        check(lib().gtb_engine_position(self.h, C.byref(p)))
        return p.value

    def read_tokens(self, first: int, count: int) -> np.ndarray:
        out = np.empty(count, np.int32)
        check(lib().gtb_engine_read_tokens(self.h, _hp(out), first, count))
        return out

    def read_logits(self) -> np.ndarray:
        out = np.empty(self.cfg.n_vocab, np.float32)
        check(lib().gtb_engine_read_logits(self.h, _hp(out)))
        return out

    def topk(self, k: int):
        """(values, ids) of the k largest logits of the last processed row, largest first, ties to the lower id."""
        v = np.empty(k, np.float32)
        ids = np.empty(k, np.int32)
        check(lib().gtb_engine_topk(self.h, k, _hp(v), _hp(ids)))
        return v, ids

    def acv(self, layer: int, aid: int) -> np.ndarray:
        out = np.empty(max(self.cfg.n_ffn, self.cfg.n_embd), np.float32)
        w = C.c_int()
        check(lib().gtb_engine_acv(self.h, layer, aid, _hp(out), C.byref(w)))
        return out[: w.value].copy()

    # ---- batched decode (gtb_engine_batch_*): up to 16 sequences share every weight read
    def batch_create(self, n_seq: int):
        check(lib().gtb_engine_batch_create(self.h, n_seq))

    def batch_prefill(self, seq: int, tokens):
        """Exact multi-row prefill of `tokens` straight into slot `seq`; the first greedy token is appended."""
        tokens = np.ascontiguousarray(tokens, np.int32)
        check(lib().gtb_engine_batch_prefill(self.h, seq, _hp(tokens), tokens.size))

    def batch_adopt(self, seq: int):
        """Slot `seq` <- the engine's current sequence (tokens, position, K/V cache)."""
        check(lib().gtb_engine_batch_adopt(self.h, seq))

    def batch_decode(self, n_steps: int):
        check(lib().gtb_engine_batch_decode(self.h, n_steps))

    def batch_position(self, seq: int) -> int:
        p = C.c_int()
        check(lib().gtb_engine_batch_position(self.h, seq, C.byref(p)))
        return p.value

    def batch_read_tokens(self, seq: int, first: int, count: int) -> np.ndarray:
        out = np.empty(count, np.int32)
        check(lib().gtb_engine_batch_read_tokens(self.h, seq, _hp(out), first, count))
        return out

    def batch_read_logits(self, seq: int) -> np.ndarray:
        out = np.empty(self.cfg.n_vocab, np.float32)
        check(lib().gtb_engine_batch_read_logits(self.h, seq, _hp(out)))
        return out

    def uses_megakernel(self) -> bool:
        y = C.c_int()
        check(lib().gtb_engine_uses_megakernel(self.h, C.byref(y)))
        return bool(y.value)

    def read_prof(self, count: int) -> np.ndarray:
        out = np.zeros(count, np.int64)
        check(lib().gtb_engine_read_prof(self.h, _hp(out), count))
        return out

    def weight_bytes(self) -> int:
        n = C.c_size_t()
        check(lib().gtb_engine_weight_bytes(self.h, C.byref(n)))
        return n.value
