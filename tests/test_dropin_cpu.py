"""Host-side pieces of the C++ drop-in headers (include/gten/): fp16 conversions and block codecs, compiled with g++ and
compared with numpy / the plain-C oracle.  No GPU, no libgten_b200 symbols needed (types + quants only)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle

ROOT = Path(__file__).resolve().parent.parent

SRC = r"""
#include "gten/gten.h"
#include <cstdio>
#include <cstring>
#include <vector>
using namespace gten;
int main(int argc, char** argv) {
    FILE* fi = std::fopen(argv[1], "rb");
    FILE* fo = std::fopen(argv[2], "wb");
    int n = 0;
    if (std::fread(&n, 4, 1, fi) != 1) return 1;
    std::vector<float> x(n);
    if (std::fread(x.data(), 4, n, fi) != (size_t)n) return 1;
    std::vector<Float16> h(n);
    std::vector<float> back(n);
    for (int i = 0; i < n; i++) { h[i] = fp32_to_fp16(x[i]); back[i] = fp16_to_fp32(h[i]); }
    std::fwrite(h.data(), 2, n, fo);
    std::fwrite(back.data(), 4, n, fo);
    const int nb = (n + 31) / 32;
    std::vector<Q8Block> q(nb);
    q8_quantize_row(x.data(), q.data(), n);
    std::fwrite(q.data(), sizeof(Q8Block), nb, fo);
    std::vector<float> dq(n);
    gten::ops::q8_dequantize_row(q.data(), dq.data(), n);     // the reference's spelling (gten/quants.h:34: namespace ops) ...
    std::vector<float> dq2(n);
    gten::q8_dequantize_row(q.data(), dq2.data(), n);          // ... and round 1's alias
    if (std::memcmp(dq.data(), dq2.data(), 4 * n) != 0) return 2;
    float (*vdp)(const char*, Dtype, const char*, Dtype, int) = &gten::ops::vec_dot_product;   // ops.h:482: the entry point exists
    if (!vdp || !gten::ops::q8_quantize_single(1.0f, 0.5f)) return 3;
    std::fwrite(dq.data(), 4, n, fo);
    return 0;
}
"""


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("dropin")
    (d / "t.cpp").write_text(SRC)
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    from tinyllama_cpp_b200 import build
    build.build()
    libdir = ROOT / "tinyllama.cpp_b200"
    subprocess.run([cxx, "-std=c++17", "-O1", "-Wall", "-Werror", "-Wno-unused-function", f"-I{ROOT / 'include'}", str(d / "t.cpp"), "-o", str(d / "t"),
                    f"-L{libdir}", "-lgten_b200", f"-Wl,-rpath,{libdir}"], check=True)
    return d / "t"


def run(exe, x, tmp_path):
    fi, fo = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fi, "wb") as f:
        f.write(np.int32(x.size).tobytes()); f.write(x.astype(np.float32).tobytes())
    subprocess.run([str(exe), str(fi), str(fo)], check=True)
    raw = fo.read_bytes()
    n = x.size
    nb = (n + 31) // 32
    h = np.frombuffer(raw[: 2 * n], np.uint16)
    back = np.frombuffer(raw[2 * n: 6 * n], np.float32)
    q = np.frombuffer(raw[6 * n: 6 * n + 34 * nb], np.uint8)
    dq = np.frombuffer(raw[6 * n + 34 * nb:], np.float32)
    return h, back, q, dq


def test_fp16_conversions_and_q8_codec(exe, tmp_path):
    rng = np.random.default_rng(3)
    special = np.array([0.0, -0.0, 1.0, -1.0, 65504.0, 65519.9, 65520.0, 65536.0, 1e9, -1e9, np.inf, -np.inf, 5.96e-8, 2.98e-8, 2.9802322e-8,
                        2.9802326e-8, 6.1e-5, 6.0975552e-5, 1e-10, 0.33325195, 0.333251953125 + 2 ** -13, 2049.0, 2051.0], np.float32)
    x = np.concatenate([special, (rng.standard_normal(4000) * np.exp(rng.standard_normal(4000) * 4)).astype(np.float32),
                        np.ldexp(rng.integers(1, 4096, 2000).astype(np.float32), rng.integers(-30, 10, 2000)).astype(np.float32)])
    h, back, q, dq = run(exe, x, tmp_path)
    with np.errstate(over="ignore"):
        want_h = x.astype(np.float16).view(np.uint16)          # numpy narrows with round-to-nearest-even, overflow -> inf
    assert np.array_equal(h, want_h)
    assert np.array_equal(back.view(np.uint32), want_h.view(np.float16).astype(np.float32).view(np.uint32))
    # NaN maps to sign | 0x7E00
    hn, _, _, _ = run(exe, np.array([np.nan, -np.nan], np.float32), tmp_path)
    assert hn[0] & 0x7fff == 0x7e00 and hn[1] & 0x7fff == 0x7e00
    # Q8 row codec against the plain-C restatement of quants.h
    port = oracle.port()
    finite = x[np.isfinite(x) & (np.abs(x) < 1e30)][:4096 + 7]
    _, _, q2, dq2 = run(exe, finite, tmp_path)
    enc = port.encode_rows(finite.reshape(1, -1), oracle.Q8)
    assert np.array_equal(q2, enc.reshape(-1))
    assert np.array_equal(dq2.view(np.uint32), port.decode_rows(enc, oracle.Q8, finite.size).reshape(-1).view(np.uint32))
