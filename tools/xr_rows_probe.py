import sys, time, json
sys.path.insert(0, '/root/repo')
import gtb
from tinyllama_cpp_b200 import capi, weights as W
capi.init(0)
cfg = W.TINYLLAMA
eng = capi.Engine(cfg, 2048, W.Q4).load(W.synth_weights(cfg, W.Q4, seed=1))
prompt = W.synth_prompt(7, 2024, cfg.n_vocab)
for rows in (64, 256, 512, 1024):
    eng.set_option("xr_rows", rows)
    eng.prefill(prompt[:300]); capi.sync()
    t0 = time.perf_counter(); eng.prefill(prompt); capi.sync(); ms = (time.perf_counter() - t0) * 1e3
    print(rows, round(ms, 2), int(eng.read_tokens(2024, 1)[0]), flush=True)
