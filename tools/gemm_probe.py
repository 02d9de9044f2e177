"""Runs the stand-alone tcgen05 GEMM on the four prefill shapes (for `ncu --metrics gpu__time_duration.sum`)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import gtb  # noqa
from tinyllama_cpp_b200 import capi
capi.init(0)
rng = np.random.default_rng(0)
bn = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for (m, n, k) in [(2048, 2048, 5632), (2048, 2048, 2048), (2048, 11264, 2048), (2048, 2560, 2048)]:
    a = rng.standard_normal((m, k)).astype(np.float16)
    w = rng.standard_normal((n, k)).astype(np.float16)
    for _ in range(3):
        capi.pf_gemm_f32(a, w, bn)
