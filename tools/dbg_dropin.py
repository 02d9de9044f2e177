"""Debug: reference app over the drop-in headers (module / op API) vs the engine, Q8, chat-template prompt."""
import ctypes as C, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gtb  # noqa
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
LIB = ROOT / "oracle" / "_ref" / "libdropin_refapp.so"
L = C.CDLL(str(LIB)); L.dropin_generate.restype = C.c_int
cfg = W.TINYLLAMA
prompt = np.array([1, 32001, 1404, 13, 22110, 338, 8425, 28579, 29973, 32002, 29871, 13, 32001, 20255, 13], np.int32)
import tempfile
for wdt, nm in ((W.Q8, "q8"), (W.Q4, "q4")):
    wl = list(W.synth_weights(cfg, wdt, seed=1))
    d = tempfile.mkdtemp(); path = Path(d) / f"t.{nm}.gten"
    W.write_gten(path, cfg, wdt, wl)
    for p, mc in ((prompt, 21), (prompt, 64), (W.synth_prompt(11, 15, cfg.n_vocab).astype(np.int32), 21)):
        toks = np.zeros(6, np.int32); lg = np.zeros(cfg.n_vocab, np.float32)
        rc = L.dropin_generate(str(path).encode(), wdt, mc, p.ctypes.data_as(C.c_void_p), p.size, 6, toks.ctypes.data_as(C.c_void_p), lg.ctypes.data_as(C.c_void_p))
        e = capi.Engine(cfg, mc, wdt).load(wl)
        want = e.generate(p, 6)
        print(nm, mc, int(p[1]), "module path:", toks.tolist(), "engine:", want[15:].tolist(), flush=True)
        e.close()
