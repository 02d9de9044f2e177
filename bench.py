#!/usr/bin/env python
"""bench.py -- batch-1 greedy decode throughput of the tinyllama.cpp forward hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload q4|q8|f16|prefill_q8|q4_seq64] [--impl b200|reference]

One "step" = one decoded token = one pass of the hot path (TinyLlama::logits for one new row + argmax).
Default workload = BASELINE.json configs[2], the configuration north_star's target is quoted on:
TinyLlama-1.1B Q4, batch 1, greedy, context ending at the full 2048-token KV cache (prompt = 2048 - K - W
synthetic ids, then W untimed + K timed decode steps).  Random-init weights in gten format (seeded, synthetic).

value   : tokens/s with everything resident in HBM: ONE persistent cooperative launch (k_mega) runs the K steps, device-side
          argmax, timed with CUDA events on the library stream (max over ranks).
e2e     : the same K steps through the reference-facing call with HOST buffers, exactly the protocol of
          greedy_sample (tinyllama.cpp:402-434): gtb_engine_logits(tokens, n, n-1) -> 128 KB of fp32 logits back
          to the host -> host argmax -> append; H2D/D2H inside the timed region.
N > 1   : independent replicas (one process per GPU, disjoint sequences, no collective on the data path);
          torch.distributed is used only for the start barrier and the max-over-ranks reduction.
seq64   : (Q4 workload, every N) BASELINE.json configs[4] in the same line: 64 sequences, sequence s on GPU s mod N, each GPU
          serves its share TOGETHER through the order-exact batched path (gtb_engine_batch_*: bit-identical tokens).
--impl reference : the reference's own CPU implementation (oracle/_ref, the unmodified sources built with
          -O3 -fopenmp -mavx -mf16c; the plain-C port if that library is absent) on all host threads, on the SAME
          configuration: the same prompt is prefilled by the reference and the same sequence positions are timed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
import gtb  # noqa: E402,F401
from tinyllama_cpp_b200 import weights as W  # noqa: E402

PREFILL_DESC = "TinyLlama-1.1B Q8 (-q8) prefill of a 2048-token synthetic prompt (BASELINE.json configs[3])"
WORKLOADS = {
    # name: (wdtype, BASELINE.json config index, description)
    "q4": (W.Q4, 2, "TinyLlama-1.1B Q4 (-q4) batch-1 greedy decode ending at the full 2048-token KV cache (BASELINE.json configs[2])"),
    "q8": (W.Q8, 1, "TinyLlama-1.1B Q8 (-q8) batch-1 greedy decode after a 128-token prompt (BASELINE.json configs[1])"),
    "f16": (W.F16, 0, "TinyLlama-1.1B FP16 batch-1 greedy decode after a 128-token prompt (BASELINE.json configs[0])"),
}
DTYPE_STR = {W.Q4: "q4 weights x q8 activations: int32 block dots, fp32 ordered accumulation",
             W.Q8: "q8 weights x q8 activations: int32 block dots, fp32 ordered accumulation",
             W.F16: "fp16 weights x fp16 activations: fp32 ordered accumulation"}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_profile(workload: str):
    """The DRAM traffic of the kernel is NOT measured in this run (that needs ncu): `roofline.traffic` is null and this object
    points at the committed `ncu --set full` capture (profiles/roofline_traffic.json: bytes per step, one code state, one t)."""
    p = ROOT / "profiles" / "roofline_traffic.json"
    if not p.exists():
        return None
    j = json.loads(p.read_text())
    d = j.get(workload)
    if d is None:
        return None
    return {"dram_bytes_per_step": d["dram_bytes_per_step"], "static": True, "source": d.get("_source", j.get("_source"))}


def use_all_host_threads():
    """The reference parallelises one loop with OpenMP (ops.h:635-637).  torchrun exports OMP_NUM_THREADS=1, so the variable is
    SET (not defaulted) before libgomp loads, and the ICV is set again through omp_set_num_threads in case it already has."""
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        import ctypes
        g = ctypes.CDLL("libgomp.so.1")
        g.omp_set_num_threads(cores)
        return int(g.omp_get_max_threads())
    except OSError:
        return cores


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def workload_shape(workload: str, K: int, Wm: int):
    """(n_prompt, max_ctx) of a decode workload: Q4 ends at the full 2048-token KV cache (configs[2]); Q8 / FP16 follow a
    128-token prompt (configs[1], [0])."""
    if workload == "q4":
        max_ctx = 2048
        n_prompt = max_ctx + 1 - K - Wm          # the last timed row is position 2047 (t = 2048)
        if n_prompt < 16:
            n_prompt, max_ctx = 16, 16 + K + Wm - 1
    else:
        n_prompt = 128
        max_ctx = n_prompt + K + Wm - 1
    return n_prompt, max_ctx


def run_reference(args, wdt, desc):
    """The reference's CPU implementation on the host cores (rank 0 only), SAME configuration as the B200 arm: the same
    synthetic prompt is prefilled by the reference itself (tinyllama.cpp:45-61 with start_pos 0), then W warm-up and K timed
    greedy steps at the same sequence positions.  For configs[2] that is a ~2000-token CPU prefill: minutes, outside the timed
    region like the B200 arm's own prefill."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = use_all_host_threads()
    import oracle
    lib = oracle.best()
    cfg = W.TINYLLAMA
    K, Wm = args.steps, args.warmup
    n_prompt, max_ctx = workload_shape(args.workload, K, Wm)
    # the reference's multi-row attention needs ceil(n/32)*34 <= max_ctx for Q8 activations, 2n <= max_ctx for FP16
    ref_ctx = max(max_ctx, 2 * n_prompt + 64 if wdt == W.F16 else ((n_prompt + 31) // 32) * 34 + 64)
    t_setup = time.perf_counter()
    m = lib.model(cfg, ref_ctx, wdt).load(W.synth_weights(cfg, wdt, seed=1))
    toks = list(W.synth_prompt(7, n_prompt, cfg.n_vocab))
    t0 = time.perf_counter()
    lg = m.logits(np.array(toks, np.int32), 0)
    prefill_s = time.perf_counter() - t0
    toks.append(int(np.argmax(lg)))
    for _ in range(Wm - 1):                       # the prefill produced the first new token, as on the B200 arm
        lg = m.logits(np.array(toks, np.int32), len(toks) - 1)
        toks.append(int(np.argmax(lg)))
    t0 = time.perf_counter()
    for _ in range(K):
        lg = m.logits(np.array(toks, np.int32), len(toks) - 1)
        toks.append(int(np.argmax(lg)))
    dt = time.perf_counter() - t0
    v = K / dt
    sample = (f"{K} decode steps at t = {n_prompt + Wm}..{n_prompt + Wm + K - 1} after the reference's own prefill of the same {n_prompt}-token prompt "
              f"({prefill_s:.1f} s = {n_prompt / prefill_s:.1f} tok/s, untimed); same weights, same positions as the B200 arm")
    print(json.dumps({
        "impl": "reference", "metric": "decode_tokens_per_s", "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": K, "warmup": Wm, "ms_per_step": 1e3 * dt / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_STR[wdt], "data": "synthetic",
        "config": {"workload": desc, "n_prompt": n_prompt, "max_ctx": max_ctx, "seq_len_timed": [n_prompt + Wm, n_prompt + Wm + K - 1],
                   "batch": 1, "sample": sample, "reference_prefill_s": prefill_s, "omp_threads": threads,
                   "setup_s": round(time.perf_counter() - t_setup, 1)},
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": threads, "kind": lib.kind, "sample": sample},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def cpu_baseline(wdt, seconds_budget=20.0):
    """Reference CPU path on a bounded sample (rank 0, N=1 only)."""
    cores = use_all_host_threads()
    import oracle
    lib = oracle.best()
    cfg = W.TINYLLAMA
    n_prompt, n_steps = 16, 48
    m = lib.model(cfg, n_prompt + n_steps + 4, wdt).load(W.synth_weights(cfg, wdt, seed=1))
    toks = list(W.synth_prompt(7, n_prompt, cfg.n_vocab))
    lg = m.logits(np.array(toks, np.int32), 0)
    toks.append(int(np.argmax(lg)))
    t0 = time.perf_counter()
    done = 0
    for _ in range(n_steps):
        lg = m.logits(np.array(toks, np.int32), len(toks) - 1)
        toks.append(int(np.argmax(lg)))
        done += 1
        if time.perf_counter() - t0 > seconds_budget:
            break
    dt = time.perf_counter() - t0
    m.close()
    return {"value": done / dt, "unit": "tokens/s", "cores": cores, "kind": lib.kind,
            "same_config": False,
            "sample": f"{done} decode steps after a {n_prompt}-token prompt (t = {n_prompt + 1}..{n_prompt + done}), same synthetic weights: a BOUNDED "
                      "short-context sample (the reference's attention is single-threaded and O(t)); `--impl reference` times the same positions as the headline"}


def prefill_flops(cfg, T):
    """layer linears of T rows + lm_head of the last row + causal attention (QK^T and P.V)."""
    per_layer = 2 * cfg.n_embd * cfg.n_embd + 2 * cfg.kv_dim * cfg.n_embd + 3 * cfg.n_ffn * cfg.n_embd
    return (2.0 * cfg.n_layers * per_layer * T + 2.0 * cfg.n_vocab * cfg.n_embd
            + cfg.n_layers * cfg.n_heads * 2.0 * 2 * cfg.d_head * T * (T + 1) / 2)


def prefill_section(capi, torch, stream, iters=5, warmup=3, T=2048, cpu=True):
    """BASELINE.json configs[3]: Q8 prefill of a 2048-token synthetic prompt through the batched tcgen05 path
    (gtb_engine_prefill_fast).  Token ids start on the host in both numbers; e2e also reads the logits back."""
    cfg = W.TINYLLAMA
    eng = capi.Engine(cfg, T + 128, W.Q8).load(W.synth_weights(cfg, W.Q8, seed=1))
    prompt = W.synth_prompt(7, T, cfg.n_vocab)
    for _ in range(warmup):
        eng.prefill_fast(prompt)
    capi.sync()
    l0 = capi.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record(stream)
    for i in range(iters):
        eng.prefill_fast(prompt)
        ev[i + 1].record(stream)
    capi.sync()
    torch.cuda.synchronize()
    ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]))
    launches = (capi.launch_count() - l0) // iters
    t0 = time.perf_counter()
    for _ in range(iters):
        eng.prefill_fast(prompt)
        lg = eng.read_logits()
    e2e_s = (time.perf_counter() - t0) / iters
    eng.close()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = float(peaks.get("bf16_tflops", 1590.0))
    fl = prefill_flops(cfg, T)
    out = {"metric": "prefill_tokens_per_s", "value": T / (ms * 1e-3), "unit": "tokens/s", "ms_per_prefill": ms,
           "config": {"workload": f"TinyLlama-1.1B Q8 (-q8) prefill of a {T}-token synthetic prompt (BASELINE.json configs[3])",
                      "path": "gtb_engine_prefill_fast: tcgen05/TMEM GEMMs fed by TMA + tensor-core causal GQA attention; tolerance-checked "
                              "against the reference (tests/test_prefill_gpu.py), the order-exact row-by-row path stays available"},
           "roofline": {"bound": "tensor", "achieved": fl / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                        "frac": fl / (ms * 1e-3) / 1e12 / peak, "flops_per_prefill": fl,
                        "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst)" if peaks else "fallback (B200_PROFILING.md)"},
           "e2e": {"value": T / e2e_s, "unit": "tokens/s", "h2d_bytes_per_step": T * 4, "d2h_bytes_per_step": cfg.n_vocab * 4},
           "gpu_launches": int(launches), "argmax_last_row": int(np.argmax(lg))}
    if cpu:
        cores = use_all_host_threads()
        import oracle
        lib = oracle.best()
        n = 24
        m = lib.model(cfg, 2 * n + 64, W.Q8).load(W.synth_weights(cfg, W.Q8, seed=1))
        t0 = time.perf_counter()
        m.logits(prompt[:n], 0)
        dt = time.perf_counter() - t0
        m.close()
        out["cpu_baseline"] = {"value": n / dt, "unit": "tokens/s", "cores": cores, "kind": lib.kind,
                               "sample": f"prefill of the first {n} prompt tokens (the reference's prefill is row-by-row GEMV, ops.h:632)"}
    return out


def fast_decode_section(eng, capi, torch, stream, cfg, wdt, prompt, n_prompt, Wm, K, hbm_peak, peak_src, e2e=True):
    """The same workload through the ORDER-FREE decode kernels (option fast_decode, gtb_fastdec.cuh): same operations and
    re-encode points as the reference, free summation order -> tolerance-level parity, NOT identical tokens.
    Reported next to the headline (which stays the bit-identical path); same timing protocol."""
    eng.set_option("fast_decode", 0)
    eng.prefill(prompt)                                   # K/V cache and first token by the exact path
    first = int(eng.read_tokens(n_prompt, 1)[0])
    toks = np.concatenate([prompt, [first]]).astype(np.int32)
    exact = eng.logits(toks, n_prompt)                    # row n_prompt through the exact kernels ...
    eng.set_option("fast_decode", 1)
    fast = eng.logits(toks, n_prompt)                     # ... and through the order-free kernels, same cache
    rel = float(np.linalg.norm(fast.astype(np.float64) - exact) / max(np.linalg.norm(exact.astype(np.float64)), 1e-30))
    eng.set_option("fast_decode", 0)
    eng.prefill(prompt)
    eng.set_option("fast_decode", 1)
    eng.decode(Wm - 1)
    capi.sync()
    torch.cuda.synchronize()
    l0 = capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    eng.decode(K)
    ev1.record(stream)
    capi.sync()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = capi.launch_count() - l0
    assert eng.position() == n_prompt + Wm - 1 + K
    t_mean = n_prompt + Wm + (K - 1) / 2.0
    bytes_per_tok = cfg.decode_bytes(wdt, int(round(t_mean)))
    achieved = bytes_per_tok * (K / (ms * 1e-3)) / 1e9
    out = {"metric": "decode_tokens_per_s", "value": K / (ms * 1e-3), "unit": "tokens/s", "ms_per_step": ms / K,
           "path": "order-free kernels (gtb_fastdec.cuh): 5 PDL-chained kernels per layer in a CUDA graph per token; 128-bit streaming "
                   "GEMV with warp-shuffle reductions, split-position attention with online-softmax combine",
           "parity": {"kind": "tolerance (summation order free; the reference's two own builds differ by the same amount, DESIGN.md 4.4/4.6)",
                      "logits_rel_l2_vs_exact_path_same_cache": rel, "same_top1": bool(int(np.argmax(fast)) == int(np.argmax(exact)))},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                        "traffic": None, "traffic_profile": traffic_profile(("q4" if wdt == W.Q4 else "q8") + "_fast_decode"),
                        "peak_source": peak_src, "algorithmic_bytes_per_step": bytes_per_tok,
                        "kernel": "k_fd_gemv / k_fd_attn chain (achieved = bytes of K steps / time of K steps)"},
           "gpu_launches": int(launches)}
    if e2e:
        eng.set_option("fast_decode", 0)
        eng.prefill(prompt)
        eng.set_option("fast_decode", 1)
        eng.decode(Wm - 1)
        n = n_prompt + Wm
        buf = np.zeros(cfg_max_ctx(eng) + 2, np.int32)
        buf[:n] = eng.read_tokens(0, n)
        capi.sync()
        t0 = time.perf_counter()
        for _ in range(K):
            lg = eng.logits(buf[:n], n - 1)
            buf[n] = int(np.argmax(lg))
            n += 1
        capi.sync()
        out["e2e"] = {"value": K / (time.perf_counter() - t0), "unit": "tokens/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": cfg.n_vocab * 4}
    eng.set_option("fast_decode", 0)
    # 8 and 16 sequences of the same length advancing together through the ORDER-FREE batch kernels (option batch_exact = 0; the
    # default batch path is the exact one, see `seq64` / `exact_batch`): every weight read is shared
    eng.set_option("batch_exact", 0)
    for B in (8, 16):
        eng.batch_create(B)
        for s in range(B):
            eng.prefill_fast(W.synth_prompt(7 + s, n_prompt, cfg.n_vocab))          # contexts by the batched tensor-core prefill
            eng.batch_adopt(s)
        eng.batch_decode(Wm)
        capi.sync()
        torch.cuda.synchronize()
        l0 = capi.launch_count()
        ev0.record(stream)
        eng.batch_decode(K - 1)
        ev1.record(stream)
        capi.sync()
        torch.cuda.synchronize()
        bms = ev0.elapsed_time(ev1)
        bbytes = cfg.weight_bytes_per_token(wdt) + B * (bytes_per_tok - cfg.weight_bytes_per_token(wdt))      # weights once + every sequence's K/V
        out["batched" if B == 8 else f"batched{B}"] = {
            "batch": B, "value": B * (K - 1) / (bms * 1e-3), "unit": "tokens/s", "ms_per_step": bms / (K - 1),
            "algorithmic_bytes_per_step": bbytes, "achieved_gbs": bbytes * (K - 1) / (bms * 1e-3) / 1e9,
            "gpu_launches": int(capi.launch_count() - l0),
            "parity": "bit-identical to the same sequences decoded one at a time with fast_decode (tests/test_fastdec_gpu.py)"}
    eng.batch_create(0)
    eng.set_option("batch_exact", 1)
    return out


def exact_batch_section(eng, capi, torch, stream, cfg, wdt, prompt, n_prompt, Wm, K):
    """B sequences at the headline's context (t ~ 2030 for configs[2]) advancing together through the ORDER-EXACT multi-row
    kernels (gtb_xrows.cu): every weight block is loaded once per step for all of them and every sequence's tokens are
    bit-identical to the headline's (checked here: all slots hold the same sequence, so they must reproduce its tokens)."""
    out = {}
    eng.prefill(prompt)
    eng.decode(Wm - 1 + K)
    want = eng.read_tokens(0, n_prompt + Wm + K)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for B in (8, 64):
        eng.batch_create(B)
        eng.prefill(prompt)
        for s in range(B):
            eng.batch_adopt(s)
        eng.batch_decode(Wm - 1)
        capi.sync()
        torch.cuda.synchronize()
        l0 = capi.launch_count()
        ev0.record(stream)
        eng.batch_decode(K)
        ev1.record(stream)
        capi.sync()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        same = all(np.array_equal(eng.batch_read_tokens(s, 0, n_prompt + Wm + K), want) for s in (0, B - 1))
        t_mean = n_prompt + Wm + (K - 1) / 2.0
        kvb = cfg.kv_bytes_per_pos(wdt) * int(round(t_mean))
        bbytes = cfg.weight_bytes_per_token(wdt) + cfg.norm_bytes_per_token() + B * kvb
        out[f"batch{B}"] = {"batch": B, "value": B * K / (ms * 1e-3), "unit": "tokens/s", "ms_per_step": ms / K,
                            "tokens_identical_to_headline": bool(same), "algorithmic_bytes_per_step": bbytes,
                            "achieved_gbs": bbytes * K / (ms * 1e-3) / 1e9, "gpu_launches": int(capi.launch_count() - l0)}
    eng.batch_create(0)
    out["path"] = ("gtb_engine_batch_decode, exact mode: k_xr_gemm (dp4a lane sums + ordered fp32 chains, R rows per weight load), "
                   "k_xr_attn, k_xr_norm; bit-identical to the reference (tests/test_xrows_gpu.py)")
    return out


def seq64_section(eng, capi, torch, stream, env, R, cfg, n_new=64, n_seq=64, n_prompt=128):
    """BASELINE.json configs[4]: 64 independent Q4 sequences (128-token prompt + n_new tokens), sequence s on GPU s mod N; each
    GPU serves its share TOGETHER: exact multi-row prefill of every prompt into its slot, then n_new - 1 steps of the exact
    batched decode (gtb_engine_batch_*).  Tokens are bit-identical to the reference's (tests/test_xrows_gpu.py checks 4 of these
    sequences at full size against committed golden tokens).  Strong scaling: the job is fixed, the GPUs share it."""
    rank, world = env.rank, env.world
    mine = R.sequences_of_rank(n_seq, rank, world)
    B = len(mine)
    prompts = [W.synth_prompt(100 + sidx, n_prompt, cfg.n_vocab) for sidx in mine]
    eng.batch_create(B)
    eng.batch_prefill(0, prompts[0])
    for s in range(1, B):
        eng.batch_prefill(s, prompts[s][:8])
    eng.batch_decode(2)                               # graph capture + warm-up
    capi.sync()
    torch.cuda.synchronize()
    R.barrier(env)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    l0 = capi.launch_count()
    ev[0].record(stream)
    for s, p in enumerate(prompts):
        eng.batch_prefill(s, p)
    ev[1].record(stream)
    eng.batch_decode(n_new - 1)
    ev[2].record(stream)
    capi.sync()
    torch.cuda.synchronize()
    pf_ms, dec_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    launches = capi.launch_count() - l0
    # end to end: prompts from host memory, every generated token back on the host, wall clock
    R.barrier(env)
    t0 = time.perf_counter()
    for s, p in enumerate(prompts):
        eng.batch_prefill(s, p)
    eng.batch_decode(n_new - 1)
    toks = [eng.batch_read_tokens(s, n_prompt, n_new) for s in range(B)]
    wall_ms = (time.perf_counter() - t0) * 1e3
    checksum = 0
    for t in toks:
        for v in t.tolist():
            checksum = (checksum * 31 + int(v)) % (1 << 31)
    eng.batch_create(0)
    dec_max, n_dec, (pf_max, all_max, wall_max), (launches,) = R.aggregate(
        env, dec_ms, B * (n_new - 1), extra_max=[pf_ms, pf_ms + dec_ms, wall_ms], extra_sum=[launches], device=f"cuda:{env.local_rank}")
    n_all = n_seq * n_new
    return {"metric": "decode_tokens_per_s", "value": R.throughput(n_dec, dec_max), "unit": "tokens/s", "scaling": "strong",
            "n_gpus": world, "sequences": n_seq, "sequences_per_gpu": B, "n_prompt": n_prompt, "n_new": n_new,
            "ms_per_step": dec_max / (n_new - 1), "prefill_ms": pf_max, "prefill_tokens_per_s": n_seq * n_prompt / (pf_max * 1e-3),
            "value_incl_prefill": n_all / (all_max * 1e-3),
            "e2e": {"value": n_all / (wall_max * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": 4.0 * n_prompt / n_new,
                    "d2h_bytes_per_step": 4.0, "note": "wall clock of the whole job on the slowest replica: prompt upload, exact prefill, decode, tokens read back"},
            "workload": "TinyLlama-1.1B Q4 decode, 64 independent sequences over the GPUs (BASELINE.json configs[4])",
            "path": "exact multi-row prefill + exact batched decode (gtb_xrows.cu): tokens bit-identical to the reference",
            "gpu_launches": int(launches), "token_checksum_rank0": checksum}


def cfg_max_ctx(eng):
    return eng.max_ctx


def run_prefill(args):
    """`--workload prefill_q8`: the whole line is BASELINE.json configs[3]; a "step" is one 2048-token prefill (N = 1 only:
    the prompt is one sequence).  `--impl reference` times the reference CPU build on a bounded sample of the same prompt."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg = W.TINYLLAMA
    T = 2048
    if args.impl == "reference":
        cores = use_all_host_threads()
        import oracle
        lib = oracle.best()
        n = 32
        m = lib.model(cfg, 2 * n + 64, W.Q8).load(W.synth_weights(cfg, W.Q8, seed=1))
        prompt = W.synth_prompt(7, T, cfg.n_vocab)
        steps = max(1, min(args.steps, 4))
        t0 = time.perf_counter()
        for _ in range(steps):
            m.logits(prompt[:n], 0)
        dt = (time.perf_counter() - t0) / steps
        v = n / dt
        sample = f"prefill of the first {n} tokens of the prompt, {steps} times (row-by-row GEMV, ops.h:632: cost per row is flat in the prompt length at this size)"
        print(json.dumps({"impl": "reference", "metric": "prefill_tokens_per_s", "value": v, "unit": "tokens/s", "n_gpus": args.gpus, "steps": steps,
                          "warmup": 0, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": DTYPE_STR[W.Q8], "data": "synthetic", "config": {"workload": PREFILL_DESC, "sample": sample},
                          "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": cores, "kind": lib.kind, "sample": sample},
                          "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    import torch
    from tinyllama_cpp_b200 import capi
    capi.init(0)
    torch.cuda.set_device(0)
    stream = torch.cuda.ExternalStream(capi.stream_handle(), device=torch.device("cuda", 0))
    with ClockSampler(0) as clk:
        p = prefill_section(capi, torch, stream, iters=max(args.steps if args.steps != 512 else 8, 3), warmup=args.warmup, T=T,
                            cpu=not args.no_cpu_baseline)
    out = {"metric": p["metric"], "value": p["value"], "unit": p["unit"], "n_gpus": 1, "steps": max(args.steps if args.steps != 512 else 8, 3),
           "warmup": args.warmup, "ms_per_step": p["ms_per_prefill"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "fp16 operands (dequantised q8 blocks) x fp16, fp32 accumulate in TMEM; q8 re-encode in the epilogues", "data": "synthetic",
           "config": dict(p["config"], l2="operands larger than L2: 2.07 GB of fp16 weight copies per prefill"),
           "roofline": dict(p["roofline"], traffic=None, kernel="k_pf_gemm<256,*> (58 % of the step; the roofline is quoted on the whole prefill)"),
           "e2e": p["e2e"], "gpu_launches": p["gpu_launches"], "clocks": clk.summary()}
    if "cpu_baseline" in p:
        out["cpu_baseline"] = p["cpu_baseline"]
    print(json.dumps(out), flush=True)


def run_seq64(args):
    """`--workload q4_seq64` = BASELINE.json configs[4] as the whole line: 64 independent sequences (128-token prompt + K new tokens
    each, K = --steps, at most 256) over N GPUs, sequence s on GPU s mod N; every GPU serves its share together through the
    order-exact batched path (seq64_section).  `one_at_a_time` repeats the job the round-1 way (batch 1 per GPU, k_mega)."""
    import torch
    from tinyllama_cpp_b200 import capi, replicas as R
    env = R.ReplicaEnv.from_env()
    rank, world, local = env.rank, env.world, env.local_rank
    if world > 1:
        torch.cuda.set_device(local)
        R.init(env, "nccl", device_id=torch.device("cuda", local))
    capi.init(local)
    torch.cuda.set_device(local)
    cfg = W.TINYLLAMA
    n_seq, n_prompt = 64, 128
    n_new = max(2, min(args.steps, 256))
    eng = capi.Engine(cfg, n_prompt + n_new, W.Q4).load(W.synth_weights(cfg, W.Q4, seed=1))
    stream = torch.cuda.ExternalStream(capi.stream_handle(), device=torch.device("cuda", local))
    with ClockSampler(local) as clk:
        sec = seq64_section(eng, capi, torch, stream, env, R, cfg, n_new=n_new, n_seq=n_seq, n_prompt=n_prompt)
    one = None
    if not args.no_fast:
        mine = R.sequences_of_rank(n_seq, rank, world)
        ms_local, toks_local = 0.0, 0
        eng.prefill(W.synth_prompt(1000, n_prompt, cfg.n_vocab))
        eng.decode(3)
        capi.sync()
        R.barrier(env)
        for sidx in mine:
            eng.prefill(W.synth_prompt(100 + sidx, n_prompt, cfg.n_vocab))
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            eng.decode(n_new - 1)
            ev1.record(stream)
            capi.sync()
            ms_local += ev0.elapsed_time(ev1)
            toks_local += n_new - 1
        ms1, units1, _, _ = R.aggregate(env, ms_local, toks_local, device=f"cuda:{local}")
        one = {"value": R.throughput(units1, ms1), "unit": "tokens/s", "path": "k_mega, one sequence at a time per GPU (round 1's way)"}
    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        B = sec["sequences_per_gpu"]
        t_mean = n_prompt + n_new / 2.0
        bytes_per_step = (cfg.weight_bytes_per_token(W.Q4) + cfg.norm_bytes_per_token() + B * cfg.kv_bytes_per_pos(W.Q4) * int(round(t_mean)))
        achieved = bytes_per_step / (sec["ms_per_step"] * 1e-3) / 1e9
        print(json.dumps({
            "metric": "decode_tokens_per_s", "value": sec["value"], "unit": "tokens/s", "n_gpus": world, "steps": n_new - 1,
            "warmup": 3, "ms_per_step": sec["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": DTYPE_STR[W.Q4], "data": "synthetic",
            "config": {"workload": sec["workload"] + f": 128-token prompt + {n_new} new tokens each, sequence s on GPU s mod N, every GPU's share decoded together (exact batched path)",
                       "sequences": n_seq, "sequences_per_gpu": B, "n_prompt": n_prompt, "n_new": n_new, "replicas": world,
                       "l2": "inputs larger than L2: every step streams 582 MB of weights"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_step": bytes_per_step,
                         "kernel": "k_xr_gemm chain of one batched step (weights once + every sequence's K/V); the step is ALU-bound (dp4a + ordered fp32 chains), not HBM-bound"},
            "e2e": sec["e2e"], "prefill_ms": sec["prefill_ms"], "prefill_tokens_per_s": sec["prefill_tokens_per_s"],
            "value_incl_prefill": sec["value_incl_prefill"], "one_at_a_time": one,
            "gpu_launches": sec["gpu_launches"], "clocks": clk.summary(), "token_checksum_rank0": sec["token_checksum_rank0"]}), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--workload", choices=list(WORKLOADS) + ["prefill_q8", "q4_seq64"], default="q4")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-prefill", action="store_true", help="skip the configs[3] prefill measurement appended to the N=1 line")
    ap.add_argument("--no-fast", action="store_true", help="skip the order-free decode measurement appended to the N=1 line")
    ap.add_argument("--no-seq64", action="store_true", help="skip the configs[4] (64 sequences) measurement appended to the Q4 line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload == "prefill_q8":
        return run_prefill(args)
    if args.workload == "q4_seq64":
        return run_seq64(args)
    wdt, cfg_idx, desc = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wdt, desc)

    import torch
    from tinyllama_cpp_b200 import capi, replicas as R
    env = R.ReplicaEnv.from_env()
    rank, world, local = env.rank, env.world, env.local_rank
    if world > 1:
        torch.cuda.set_device(local)
        R.init(env, "nccl", device_id=torch.device("cuda", local))
    capi.init(local)
    torch.cuda.set_device(local)
    cfg = W.TINYLLAMA
    K, Wm = args.steps, args.warmup
    n_prompt, max_ctx = workload_shape(args.workload, K, Wm)
    t_setup = time.time()
    eng = capi.Engine(cfg, max_ctx, wdt).load(W.synth_weights(cfg, wdt, seed=1))
    assert eng.uses_megakernel(), "bench must run the persistent-kernel path"
    prompt = W.synth_prompt(7 + rank, n_prompt, cfg.n_vocab)       # disjoint sequences per replica
    stream = torch.cuda.ExternalStream(capi.stream_handle(), device=torch.device("cuda", local))
    hbm_peak, peak_src = measured_peaks()

    def barrier():
        capi.sync()
        torch.cuda.synchronize()
        R.barrier(env)

    # ---- resident path: exact prefill (outside the headline's timed region; timed on its own), W warm-up steps, K timed steps
    eng.prefill(prompt[:min(600, n_prompt)])      # warm-up: one full 512-row pass + a partial one loads every kernel variant the timed prefill uses
    barrier()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record(stream)
    eng.prefill(prompt)
    pe1.record(stream)
    capi.sync()
    torch.cuda.synchronize()
    prefill_ms = pe0.elapsed_time(pe1)
    eng.decode(Wm - 1)            # prefill already produced the first new token
    barrier()
    l0 = capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record(stream)
        eng.decode(K)
        ev1.record(stream)
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = capi.launch_count() - l0
    pos_end = eng.position()
    assert pos_end == n_prompt + Wm - 1 + K, (pos_end, n_prompt, Wm, K)
    toks_resident = eng.read_tokens(0, pos_end + 1)

    # ---- e2e path: host buffers, the reference's per-token protocol
    e2e = None
    if not args.no_e2e:
        eng.prefill(prompt)
        eng.decode(Wm - 1)
        toks = eng.read_tokens(0, n_prompt + Wm).astype(np.int32)
        buf = np.zeros(max_ctx + 2, np.int32)
        n = toks.size
        buf[:n] = toks
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            lg = eng.logits(buf[:n], n - 1)                 # 4 B H2D + n_vocab*4 B D2H inside
            buf[n] = int(np.argmax(lg))                      # first maximum wins, like tinyllama.cpp:416-424
            n += 1
        capi.sync()
        e2e_s = time.perf_counter() - t0
        assert np.array_equal(buf[:n], toks_resident[:n]), "e2e and resident greedy sequences differ"
        e2e_local = e2e_s
    else:
        e2e_local = float("nan")

    # replicas: the job is done when the slowest replica is; units (tokens) and launches add up
    ms, units, (e2e_ms,), (launches,) = R.aggregate(env, ms, K, extra_max=[e2e_local * 1e3], extra_sum=[launches], device=f"cuda:{local}")
    launches = int(launches)
    value = R.throughput(units, ms)
    t_mean = n_prompt + Wm + (K - 1) / 2.0                 # mean sequence length over the timed steps
    bytes_per_tok = cfg.decode_bytes(wdt, int(round(t_mean)))
    achieved = bytes_per_tok * (K / (ms * 1e-3)) / 1e9     # per GPU
    clocks = clk.summary()

    if rank == 0:
        out = {
            "metric": "decode_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE_STR[wdt], "data": "synthetic",
            "config": {"workload": desc, "baseline_config_index": cfg_idx, "n_prompt": n_prompt, "max_ctx": max_ctx,
                       "seq_len_timed": [n_prompt + Wm, n_prompt + Wm + K - 1], "batch": 1, "replicas": world,
                       "weights": "random-init, seeded, gten format", "l2": "inputs larger than L2: every step streams "
                       f"{cfg.weight_bytes_per_token(wdt) / 1e6:.0f} MB of weights (L2 = 126 MB)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": None, "traffic_profile": traffic_profile(args.workload), "peak_source": peak_src,
                         "kernel": "k_mega<%s>: one persistent cooperative launch runs all K steps (achieved = bytes of K steps / launch time)" % args.workload,
                         "algorithmic_bytes_per_step": bytes_per_tok},
            "gpu_launches": launches, "clocks": clocks,
        }
        if not args.no_e2e:
            out["e2e"] = {"value": units / (e2e_ms * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": 4,
                          "d2h_bytes_per_step": cfg.n_vocab * 4}
        out["exact_prefill"] = {"tokens": n_prompt, "ms": prefill_ms, "tokens_per_s": n_prompt / (prefill_ms * 1e-3),
                                "path": "order-exact multi-row kernels (gtb_xrows.cu), 512 rows per pass: bit-identical to the reference's row loop"}
        if world == 1 and not args.no_fast and wdt != W.F16:
            out["fast_decode"] = fast_decode_section(eng, capi, torch, stream, cfg, wdt, prompt, n_prompt, Wm, K, hbm_peak, peak_src,
                                                     e2e=not args.no_e2e)
        if world == 1 and not args.no_fast:
            out["exact_batch"] = exact_batch_section(eng, capi, torch, stream, cfg, wdt, prompt, n_prompt, Wm, K)
    # configs[4] at this N (every rank takes part: strong scaling over the replicas)
    seq64 = None
    if args.workload == "q4" and not args.no_seq64:
        seq64 = seq64_section(eng, capi, torch, stream, env, R, cfg)
    if rank == 0:
        if seq64 is not None:
            out["seq64"] = seq64
        if world == 1 and not args.no_prefill:
            eng.close()
            out["prefill"] = prefill_section(capi, torch, stream, cpu=not args.no_cpu_baseline)
        if world == 1 and not args.no_cpu_baseline:
            eng.close()
            out["cpu_baseline"] = cpu_baseline(wdt)
        out["setup_s"] = round(time.time() - t_setup, 1)
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
