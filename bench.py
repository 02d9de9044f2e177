#!/usr/bin/env python
"""bench.py -- batch-1 greedy decode throughput of the tinyllama.cpp forward hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload q4|q8|f16|prefill_q8|q4_seq64] [--impl b200|reference]

One "step" = one decoded token = one pass of the hot path (TinyLlama::logits for one new row + argmax).
Default workload = BASELINE.json configs[2], the configuration north_star's target is quoted on:
TinyLlama-1.1B Q4, batch 1, greedy, context ending at the full 2048-token KV cache (prompt = 2048 - K - W
synthetic ids, then W untimed + K timed decode steps).  Random-init weights in gten format (seeded, synthetic).

value   : tokens/s with everything resident in HBM: K CUDA-graph replays back to back, device-side argmax,
          timed with CUDA events on the library stream (max over ranks).
e2e     : the same K steps through the reference-facing call with HOST buffers, exactly the protocol of
          greedy_sample (tinyllama.cpp:402-434): gtb_engine_logits(tokens, n, n-1) -> 128 KB of fp32 logits back
          to the host -> host argmax -> append; H2D/D2H inside the timed region.
N > 1   : independent replicas (one process per GPU, disjoint sequences, no collective on the data path);
          torch.distributed is used only for the start barrier and the max-over-ranks reduction.
--impl reference : the reference's own CPU implementation (oracle/_ref, the unmodified sources built with
          -O3 -fopenmp -mavx -mf16c; the plain-C port if that library is absent) on all host threads.
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


def measured_traffic(workload: str, steps: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the kernel from the committed `ncu --set full` capture, scaled to the
    K steps one launch of this run processes (profiles/roofline_traffic.json: bytes per step), or None."""
    p = ROOT / "profiles" / "roofline_traffic.json"
    if not p.exists():
        return None
    d = json.loads(p.read_text()).get(workload)
    return None if d is None else d["dram_bytes_per_step"] * steps


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


def run_reference(args, wdt, desc):
    """The reference's CPU implementation on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import oracle
    lib = oracle.best()
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    cfg = W.TINYLLAMA
    n_prompt = 32
    max_ctx = n_prompt + args.warmup + args.steps + 2
    m = lib.model(cfg, max_ctx, wdt).load(W.synth_weights(cfg, wdt, seed=1))
    toks = list(W.synth_prompt(7, n_prompt, cfg.n_vocab))
    lg = m.logits(np.array(toks, np.int32), 0)
    toks.append(int(np.argmax(lg)))
    for _ in range(args.warmup):
        lg = m.logits(np.array(toks, np.int32), len(toks) - 1)
        toks.append(int(np.argmax(lg)))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lg = m.logits(np.array(toks, np.int32), len(toks) - 1)
        toks.append(int(np.argmax(lg)))
    dt = time.perf_counter() - t0
    v = args.steps / dt
    sample = f"{args.steps} decode steps after a {n_prompt}-token prompt (t = {n_prompt + args.warmup + 1}..{len(toks) - 1}); short context: the reference's attention is single-threaded"
    print(json.dumps({
        "impl": "reference", "metric": "decode_tokens_per_s", "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_STR[wdt], "data": "synthetic",
        "config": {"workload": desc, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": cores, "kind": lib.kind, "sample": sample},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def cpu_baseline(wdt, seconds_budget=20.0):
    """Reference CPU path on a bounded sample (rank 0, N=1 only)."""
    import oracle
    lib = oracle.best()
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
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
            "sample": f"{done} decode steps after a {n_prompt}-token prompt (t = {n_prompt + 1}..{n_prompt + done}), same synthetic weights"}


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
        import oracle
        lib = oracle.best()
        n = 24
        m = lib.model(cfg, 2 * n + 64, W.Q8).load(W.synth_weights(cfg, W.Q8, seed=1))
        t0 = time.perf_counter()
        m.logits(prompt[:n], 0)
        dt = time.perf_counter() - t0
        m.close()
        out["cpu_baseline"] = {"value": n / dt, "unit": "tokens/s", "cores": os.cpu_count() or 1, "kind": lib.kind,
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
                        "traffic": measured_traffic(("q4" if wdt == W.Q4 else "q8") + "_fast_decode", K),
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
    # 8 and 16 sequences of the same length advancing together (gtb_engine_batch_*): every weight read is shared
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
    return out


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
        import oracle
        lib = oracle.best()
        cores = os.cpu_count() or 1
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
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
    """`--workload q4_seq64` = BASELINE.json configs[4]: 64 independent sequences (128-token prompt + K new tokens each, K = --steps,
    at most 256) served by N replicas, sequence s on GPU s mod N, back to back at batch 1 (the reference has no batch dimension,
    SURVEY.md 8e).  value = generated tokens of all sequences / the slowest replica's summed decode time (CUDA events around each
    sequence's device-resident decode; the exact prefill of the prompt is outside the timed region)."""
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
    mine = R.sequences_of_rank(n_seq, rank, world)
    eng.prefill(W.synth_prompt(1000, n_prompt, cfg.n_vocab))        # warm-up sequence (not counted)
    eng.decode(max(args.warmup, 3))
    capi.sync()
    R.barrier(env)
    ms_local, toks_local, checksum = 0.0, 0, 0
    l0 = capi.launch_count()
    t_wall = time.perf_counter()
    with ClockSampler(local) as clk:
        for sidx in mine:
            eng.prefill(W.synth_prompt(100 + sidx, n_prompt, cfg.n_vocab))      # produces the first new token
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            eng.decode(n_new - 1)
            ev1.record(stream)
            capi.sync()
            ms_local += ev0.elapsed_time(ev1)
            toks_local += n_new - 1
            checksum = (checksum * 31 + int(eng.read_tokens(n_prompt + n_new - 1, 1)[0])) % (1 << 31)
    wall = time.perf_counter() - t_wall
    launches = capi.launch_count() - l0
    ms, units, (wall_max,), (launches,) = R.aggregate(env, ms_local, toks_local, extra_max=[wall], extra_sum=[launches], device=f"cuda:{local}")
    # the same job through the batched order-free decode (gtb_engine_batch_*): this replica's sequences in groups
    # of up to 16 that share every weight read; tolerance-level parity, reported next to the bit-exact number
    BATCH = 16
    bms_local, btoks_local, bl0 = 0.0, 0, capi.launch_count()
    if not args.no_fast:
        for g0 in range(0, len(mine), BATCH):
            grp = mine[g0:g0 + BATCH]
            eng.batch_create(len(grp))
            for slot, sidx in enumerate(grp):
                eng.prefill(W.synth_prompt(100 + sidx, n_prompt, cfg.n_vocab))
                eng.batch_adopt(slot)
            if g0 == 0:
                eng.batch_decode(1)           # graph capture outside the timed region; re-adopt so that every sequence starts at n_prompt
                for slot, sidx in enumerate(grp):
                    eng.prefill(W.synth_prompt(100 + sidx, n_prompt, cfg.n_vocab))
                    eng.batch_adopt(slot)
            capi.sync()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            eng.batch_decode(n_new - 1)
            ev1.record(stream)
            capi.sync()
            bms_local += ev0.elapsed_time(ev1)
            btoks_local += len(grp) * (n_new - 1)
            assert all(eng.batch_position(slot) == n_prompt + n_new - 1 for slot in range(len(grp)))
        eng.batch_create(0)
    blaunches = capi.launch_count() - bl0
    bms, bunits, _, (blaunches,) = R.aggregate(env, bms_local, btoks_local, extra_max=[0.0], extra_sum=[blaunches], device=f"cuda:{local}")
    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        t_mean = n_prompt + n_new / 2.0
        bytes_per_tok = cfg.decode_bytes(W.Q4, int(round(t_mean)))
        per_gpu_tok_s = (len(mine) * (n_new - 1)) / (ms_local * 1e-3) if ms_local else 0.0
        achieved = bytes_per_tok * per_gpu_tok_s / 1e9
        print(json.dumps({
            "metric": "decode_tokens_per_s", "value": R.throughput(units, ms), "unit": "tokens/s", "n_gpus": world, "steps": n_new - 1,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / max(1, len(mine) * (n_new - 1)), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": DTYPE_STR[W.Q4], "data": "synthetic",
            "config": {"workload": "TinyLlama-1.1B Q4 decode, 64 independent sequences as replicas (BASELINE.json configs[4]): 128-token prompt + "
                                   f"{n_new} new tokens each, sequence s on GPU s mod N, batch 1 per GPU", "sequences": n_seq,
                       "sequences_per_gpu": len(mine), "n_prompt": n_prompt, "n_new": n_new, "replicas": world,
                       "l2": "inputs larger than L2: every step streams 582 MB of weights"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_step": bytes_per_tok, "kernel": "k_mega<q4>, one launch per sequence"},
            "e2e": {"value": units / wall_max if wall_max else None, "unit": "tokens/s", "h2d_bytes_per_step": 4 * n_prompt / (n_new - 1),
                    "d2h_bytes_per_step": 4.0 / (n_new - 1), "note": "wall clock of the whole job on the slowest replica, prompt upload and exact prefill included"},
            "batched": None if args.no_fast or not bms else {
                "value": R.throughput(bunits, bms), "unit": "tokens/s", "batch_per_gpu": min(BATCH, len(mine)),
                "ms_per_step": bms / max(1, -(-len(mine) // BATCH) * (n_new - 1)),
                "path": "gtb_engine_batch_decode: order-free kernels, up to 16 sequences share every weight read; a sequence decoded in a batch "
                        "gives the same bits as decoded alone with fast_decode (tests/test_fastdec_gpu.py); tolerance-level parity with the reference",
                "gpu_launches": int(blaunches)},
            "gpu_launches": int(launches), "clocks": clk.summary(), "token_checksum_rank0": checksum}), flush=True)
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
    if args.workload == "q4":
        max_ctx = 2048
        n_prompt = max_ctx + 1 - K - Wm          # the last timed row is position 2047 (t = 2048)
        if n_prompt < 16:
            n_prompt, max_ctx = 16, 16 + K + Wm - 1
    else:
        n_prompt = 128
        max_ctx = n_prompt + K + Wm - 1
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

    # ---- resident path: prefill (untimed, exact row-by-row path), W warm-up steps, K timed steps
    eng.prefill(prompt)
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
                         "traffic": measured_traffic(args.workload, K), "peak_source": peak_src,
                         "kernel": "k_mega<%s>: one persistent cooperative launch runs all K steps (achieved = bytes of K steps / launch time)" % args.workload,
                         "algorithmic_bytes_per_step": bytes_per_tok},
            "gpu_launches": launches, "clocks": clocks,
        }
        if not args.no_e2e:
            out["e2e"] = {"value": units / (e2e_ms * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": 4,
                          "d2h_bytes_per_step": cfg.n_vocab * 4}
        if world == 1 and not args.no_fast and wdt != W.F16:
            out["fast_decode"] = fast_decode_section(eng, capi, torch, stream, cfg, wdt, prompt, n_prompt, Wm, K, hbm_peak, peak_src,
                                                     e2e=not args.no_e2e)
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
