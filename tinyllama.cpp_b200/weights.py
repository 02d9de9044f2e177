"""Model shapes, synthetic gten-format weights and the .gten container.

Restates the *format* side of the reference: the weight quantisers and record writer of
tinyllama_to_gten.py (q8_quantize :24-51, q4_quantize :54-91, write_layer :94-148, tensor order
:151-201) and the reader's expectations in tinyllama.cpp:301-392.  Pure numpy; no GPU involved.

Synthetic weights are made from an integer hash (splitmix64) so that every platform produces the
same bytes without touching libm: each weight is the centred sum of the four 16-bit fields of one
64-bit hash (Irwin-Hall, n=4: bell-shaped, bounded at +-3.46 sigma) scaled to sigma = 0.02; norm
weights are 1 + 0.1 * that unit variate (the recipe of SURVEY.md §8c with the Box-Muller step
replaced by an exact-integer one).
"""
from __future__ import annotations

import struct
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from pathlib import Path
from typing import Iterator, List, Tuple

import numpy as np

I32, F16, F32, Q8, Q4 = 0, 1, 2, 3, 4            # gten_types.h:20-26
WDTYPE_NAMES = {F16: "fp16", Q8: "q8", Q4: "q4"}
WDTYPE_BY_NAME = {v: k for k, v in WDTYPE_NAMES.items()}

# tensor ids shared with include/gten_b200.h and oracle/
T_EMBED, T_FINAL_NORM, T_LM_HEAD = 0, 1, 2
T_Q, T_K, T_V, T_O, T_GATE, T_UP, T_DOWN, T_ATTN_NORM, T_FFN_NORM = 10, 11, 12, 13, 14, 15, 16, 17, 18

GTEN_MAGIC = 0x454C49464E455447                  # "GTENFILE", tinyllama.cpp:340


@dataclass(frozen=True)
class ModelConfig:
    """tinyllama.cpp:12-20 (TinyLLamaParams)."""
    n_vocab: int = 32003
    n_embd: int = 2048
    n_ffn: int = 5632
    n_layers: int = 22
    n_heads: int = 32
    n_groups: int = 4

    @property
    def d_head(self) -> int:
        return self.n_embd // self.n_heads

    @property
    def kv_dim(self) -> int:
        return self.d_head * self.n_groups

    def weight_bytes_per_token(self, wdtype: int) -> int:
        """Bytes of every matrix a decode step streams (155 GEMVs incl. lm_head), SURVEY.md §8(a11)."""
        per_layer = (2 * self.n_embd * self.n_embd + 2 * self.kv_dim * self.n_embd + 3 * self.n_ffn * self.n_embd)
        n = self.n_layers * per_layer + self.n_vocab * self.n_embd
        return {F16: n * 2, Q8: n // 32 * 34, Q4: n // 32 * 18}[wdtype]

    def norm_bytes_per_token(self) -> int:
        return (2 * self.n_layers + 1) * self.n_embd * 2

    def kv_bytes_per_pos(self, wdtype: int) -> int:
        """K and V bytes read per cached position per token (all layers)."""
        row = self.kv_dim * 2 if wdtype == F16 else self.kv_dim // 32 * 34
        return 2 * self.n_layers * row

    def decode_bytes(self, wdtype: int, t: int) -> int:
        """Algorithmic HBM bytes of one decode step at sequence length t."""
        return self.weight_bytes_per_token(wdtype) + self.norm_bytes_per_token() + self.kv_bytes_per_pos(wdtype) * t


TINYLLAMA = ModelConfig()


def mini_config(n_layers: int = 2, n_vocab: int = 512, n_ffn: int = 5632, n_embd: int = 2048,
                n_heads: int = 32, n_groups: int = 4) -> ModelConfig:
    """A reduced network (same head geometry) for fast parity tests."""
    return ModelConfig(n_vocab=n_vocab, n_embd=n_embd, n_ffn=n_ffn, n_layers=n_layers, n_heads=n_heads, n_groups=n_groups)


def row_nbytes(dtype: int, n: int) -> int:
    if dtype == Q8:
        return ((n + 31) // 32) * 34
    if dtype == Q4:
        return (n // 32) * 18
    if dtype == F16:
        return n * 2
    return n * 4


def tensor_list(cfg: ModelConfig) -> List[Tuple[int, int, str, int, int]]:
    """(layer, tensor id, checkpoint name, rows, cols) in file order (tinyllama_to_gten.py:157-201)."""
    out = [(0, T_EMBED, "model.embed_tokens.weight", cfg.n_vocab, cfg.n_embd)]
    for i in range(cfg.n_layers):
        b = f"model.layers.{i}"
        out += [
            (i, T_Q, f"{b}.self_attn.q_proj.weight", cfg.n_embd, cfg.n_embd),
            (i, T_K, f"{b}.self_attn.k_proj.weight", cfg.kv_dim, cfg.n_embd),
            (i, T_V, f"{b}.self_attn.v_proj.weight", cfg.kv_dim, cfg.n_embd),
            (i, T_O, f"{b}.self_attn.o_proj.weight", cfg.n_embd, cfg.n_embd),
            (i, T_GATE, f"{b}.mlp.gate_proj.weight", cfg.n_ffn, cfg.n_embd),
            (i, T_UP, f"{b}.mlp.up_proj.weight", cfg.n_ffn, cfg.n_embd),
            (i, T_DOWN, f"{b}.mlp.down_proj.weight", cfg.n_embd, cfg.n_ffn),
            (i, T_ATTN_NORM, f"{b}.input_layernorm.weight", 1, cfg.n_embd),
            (i, T_FFN_NORM, f"{b}.post_attention_layernorm.weight", 1, cfg.n_embd),
        ]
    out += [(0, T_FINAL_NORM, "model.norm.weight", 1, cfg.n_embd), (0, T_LM_HEAD, "lm_head.weight", cfg.n_vocab, cfg.n_embd)]
    return out


def is_norm(tid: int) -> bool:
    return tid in (T_FINAL_NORM, T_ATTN_NORM, T_FFN_NORM)


# ------------------------------------------------------------------ synthetic values ----
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_IH_STD = float(np.sqrt(4.0 * (65536.0 ** 2 - 1.0) / 12.0))      # std of the sum of four uniform 16-bit ints


def _mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on a uint64 array (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def unit_variates(seed: int, uid: int, n: int, offset: int = 0) -> np.ndarray:
    """n float32 values with mean 0 and variance 1, a pure function of (seed, uid, offset+i)."""
    with np.errstate(over="ignore"):
        base = _mix64(np.array([np.uint64(seed) * np.uint64(0x2545F4914F6CDD1D) + np.uint64(uid)], dtype=np.uint64))[0]
        z = _mix64(base + np.arange(offset, offset + n, dtype=np.uint64))
    s = (z & np.uint64(0xFFFF)).astype(np.int64)
    s += ((z >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)
    s += ((z >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int64)
    s += (z >> np.uint64(48)).astype(np.int64)
    s -= 131070
    return s.astype(np.float32) * np.float32(1.0 / _IH_STD)


def synth_tensor(seed: int, layer: int, tid: int, rows: int, cols: int, row0: int = 0, nrows: int | None = None) -> np.ndarray:
    """float32 [nrows, cols] slice of a synthetic tensor: N(0, 0.02^2)-like, or 1 + 0.1 u for norms."""
    nrows = rows - row0 if nrows is None else nrows
    uid = layer * 64 + tid
    u = unit_variates(seed, uid, nrows * cols, offset=row0 * cols).reshape(nrows, cols)
    if is_norm(tid):
        return (np.float32(1.0) + np.float32(0.1) * u).astype(np.float32)
    return (np.float32(0.02) * u).astype(np.float32)


# ------------------------------------------------------------------ converter quantisers ----
def q8_quantize(t: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """tinyllama_to_gten.py:24-51 -> (fp16 deltas [n_blocks], int8 codes [n_blocks, 32]); ties-to-even."""
    assert t.ndim == 2 and t.shape[1] % 32 == 0
    b = np.ascontiguousarray(t, dtype=np.float32).reshape(-1, 32)
    deltas = (np.abs(b).max(axis=1) / np.float32(127.0)).astype(np.float32)
    scal = deltas.copy()
    nz = scal != 0
    scal[nz] = np.float32(1.0) / scal[nz]
    q = np.rint(b * scal[:, None]).astype(np.int8)
    return deltas.astype(np.float16), q


def q4_quantize(t: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """tinyllama_to_gten.py:54-91 -> (fp16 deltas, uint8 packed [n_blocks, 16]); byte j = (elt j + 7) << 4 | (elt j+16 + 7)."""
    assert t.ndim == 2 and t.shape[1] % 32 == 0
    b = np.ascontiguousarray(t, dtype=np.float32).reshape(-1, 32)
    deltas = (np.abs(b).max(axis=1) / np.float32(7.0)).astype(np.float32)
    scal = deltas.copy()
    nz = scal != 0
    scal[nz] = np.float32(1.0) / scal[nz]
    q = (np.rint(b * scal[:, None]) + np.float32(7.0)).astype(np.uint8)
    packed = (q[:, :16] << 4) | (q[:, 16:] & 0x0F)
    return deltas.astype(np.float16), packed.astype(np.uint8)


def quantize_payload(w: np.ndarray, wdtype: int) -> np.ndarray:
    """float32 [rows, cols] -> the .gten payload bytes of that tensor (write_layer, tinyllama_to_gten.py:94-148)."""
    if wdtype == F16:
        return np.ascontiguousarray(w, dtype=np.float32).astype(np.float16).view(np.uint8).reshape(-1)
    if wdtype == Q8:
        d, q = q8_quantize(w)
        out = np.empty((d.size, 34), np.uint8)
        out[:, :2] = d.view(np.uint8).reshape(-1, 2)
        out[:, 2:] = q.view(np.uint8)
        return out.reshape(-1)
    if wdtype == Q4:
        d, q = q4_quantize(w)
        out = np.empty((d.size, 18), np.uint8)
        out[:, :2] = d.view(np.uint8).reshape(-1, 2)
        out[:, 2:] = q
        return out.reshape(-1)
    raise ValueError(wdtype)


def dequantize_payload(payload: np.ndarray, wdtype: int, rows: int, cols: int) -> np.ndarray:
    """Payload bytes -> float32 [rows, cols] exactly as quants.h:69-90 decodes them."""
    payload = np.ascontiguousarray(payload).view(np.uint8).reshape(-1)
    if wdtype == F16:
        return payload.view(np.float16).astype(np.float32).reshape(rows, cols)
    if wdtype == Q8:
        b = payload.reshape(-1, 34)
        d = b[:, :2].copy().view(np.float16).astype(np.float32)
        return (b[:, 2:].view(np.int8).astype(np.float32) * d).reshape(rows, cols)
    b = payload.reshape(-1, 18)
    d = b[:, :2].copy().view(np.float16).astype(np.float32)
    hi = (b[:, 2:] >> 4).astype(np.int32) - 7
    lo = (b[:, 2:] & 0x0F).astype(np.int32) - 7
    return (np.concatenate([hi, lo], axis=1).astype(np.float32) * d).reshape(rows, cols)


def synth_payload(seed: int, wdtype: int, layer: int, tid: int, rows: int, cols: int, chunk_rows: int = 2048) -> np.ndarray:
    dt = F16 if is_norm(tid) else wdtype          # norm weights are always fp16 (tinyllama_to_gten.py:191-198)
    parts = []
    for r0 in range(0, rows, chunk_rows):
        n = min(chunk_rows, rows - r0)
        parts.append(quantize_payload(synth_tensor(seed, layer, tid, rows, cols, r0, n), dt))
    return parts[0] if len(parts) == 1 else np.concatenate(parts)


def synth_weights(cfg: ModelConfig, wdtype: int, seed: int = 1, threads: int = 8) -> Iterator[Tuple[int, int, np.ndarray]]:
    """Yields (layer, tensor id, payload) in checkpoint order."""
    tl = tensor_list(cfg)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        futs = [ex.submit(synth_payload, seed, wdtype, l, t, r, c) for (l, t, _n, r, c) in tl]
        for (l, t, _n, _r, _c), f in zip(tl, futs):
            yield l, t, f.result()


def synth_prompt(seed: int, n: int, n_vocab: int = 32000) -> np.ndarray:
    """n uniform token ids in [0, min(n_vocab, 32000)) as int32 (prompt recipe of SURVEY.md §8c)."""
    hi = min(n_vocab, 32000)
    with np.errstate(over="ignore"):
        z = _mix64(np.uint64(seed) * np.uint64(0xD1342543DE82EF95) + np.arange(n, dtype=np.uint64) + np.uint64(0x5EED))
    return (z % np.uint64(hi)).astype(np.int32)


# ------------------------------------------------------------------ .gten container ----
def write_gten(path, cfg: ModelConfig, wdtype: int, weights) -> None:
    """weights: iterable of (layer, tid, payload) in tensor_list order.  Layout: tinyllama_to_gten.py:94-201."""
    names = {(l, t): n for (l, t, n, _r, _c) in tensor_list(cfg)}
    with open(path, "wb") as f:
        f.write(struct.pack("<q", GTEN_MAGIC))
        for layer, tid, payload in weights:
            name = names[(layer, tid)].encode()
            payload = np.ascontiguousarray(payload).view(np.uint8).reshape(-1)
            f.write(struct.pack("<i", len(name)) + name)       # layer header
            f.write(struct.pack("<i", len(name)) + name)       # weight name
            f.write(struct.pack("<i", payload.size))
            f.write(payload.tobytes())


def read_gten(path, cfg: ModelConfig, wdtype: int) -> Iterator[Tuple[int, int, np.ndarray]]:
    """Mirror of TinyLlama::load_from_ckpt (tinyllama.cpp:336-392): fixed order, size-checked payloads."""
    with open(path, "rb") as f:
        (magic,) = struct.unpack("<q", f.read(8))
        if magic != GTEN_MAGIC:
            raise ValueError("Magic number in the binary does not match the expected one.")
        for layer, tid, name, rows, cols in tensor_list(cfg):
            (n,) = struct.unpack("<i", f.read(4)); f.read(n)
            (n,) = struct.unpack("<i", f.read(4)); wname = f.read(n).decode()
            (nb,) = struct.unpack("<i", f.read(4))
            expect = rows * row_nbytes(F16 if is_norm(tid) else wdtype, cols)
            if nb != expect:
                raise ValueError(f"Weight `{wname}` data size: {nb} does not match the expected size: {expect}.")
            yield layer, tid, np.frombuffer(f.read(nb), dtype=np.uint8)
