"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol that
include/gten_b200.h declares, and refuses loudly to compute without a GPU (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def capi():
    from tinyllama_cpp_b200 import build, capi
    build.build()
    return capi


def test_header_symbols_are_exported(capi):
    header = (ROOT / "include" / "gten_b200.h").read_text()
    declared = set(re.findall(r"\b(gtb_[a-z0-9_]+)\s*\(", header))
    declared -= {"gtb_weight", "gtb_engine"}
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    L = capi.lib()
    for s in declared:
        assert hasattr(L, s), s


def test_no_cpu_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = capi.lib()
    n = C.c_int(-1)
    assert L.gtb_device_count(C.byref(n)) == 0 and n.value == 0
    assert L.gtb_init(0) == 3          # GTB_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.gtb_last_error()
    p = C.c_void_p()
    assert L.gtb_malloc(C.byref(p), 16) != 0
    with pytest.raises(capi.GtbError):
        capi.init(0)


def test_product_does_not_import_oracle():
    for f in (ROOT / "tinyllama.cpp_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h", ".cpp") and "build" not in f.parts:
            t = f.read_text()
            assert "import oracle" not in t and "from oracle" not in t and "liboracle" not in t and "libgten_ref" not in t, f
    for f in (ROOT / "include").rglob("*"):
        if f.is_file():
            assert "oracle" not in f.read_text(), f


def test_every_engine_option_is_documented_in_the_header():
    """gtb_engine_set_option names handled in gtb_engine.cu all appear in include/gten_b200.h (the C-ABI's documentation)."""
    import re
    root = Path(__file__).resolve().parent.parent
    src = (root / "tinyllama.cpp_b200" / "csrc" / "gtb_engine.cu").read_text()
    hdr = (root / "include" / "gten_b200.h").read_text()
    names = sorted(set(re.findall(r'!strcmp\(name, "([a-z_0-9]+)"\)', src)))
    assert len(names) > 20
    missing = [n for n in names if f'"{n}"' not in hdr]
    assert not missing, f"options without documentation in gten_b200.h: {missing}"
