import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import gtb  # noqa
from tinyllama_cpp_b200 import capi
capi.init(0)
rng = np.random.default_rng(0)
for n in (2048, 300):
    q = rng.integers(-127, 128, n).astype(np.float32) * np.float32(0.0123)
    t = (q * q).astype(np.float32)
    for _ in range(3):
        r, c = capi.selftest_exact_sum(t, cycles=True)
        print(n, r, c)
