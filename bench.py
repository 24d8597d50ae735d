#!/usr/bin/env python
"""bench.py — batch-1 f32 greedy decode throughput of the B200-native rama decode path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model llama2-7B] [--impl ours|reference]

A *step* is one full greedy generation of `--tokens` (default 256) tokens — BOS, the prompt
"once upon a time" forced, then argmax sampling — i.e. one pass of the reference's `generate`
loop (engine/src/transformer/mod.rs:169-206) over one synthetic input.  Weights are synthetic
(counter-based N(0,1/sqrt(D)), seed 1234, llama2.c v0 layout), generated directly in
HBM.  N > 1 (launched by torchrun, one process per GPU) runs the same model tensor-parallel over
NCCL: strong scaling.

Prints ONE JSON line (rank 0).  `value` = device-resident loop (token feedback on the device,
CUDA-event time, max over ranks).  `e2e` = the reference-facing call sequence
forward(token,pos) + sample() per token with the token id crossing PCIe both ways every step.
`roofline` = the dominant kernel (fused rmsnorm→[w1|w3]→SwiGLU GEMV) against the measured HBM peak.
`cpu_baseline` = the CPU oracle (C++ restatement of the reference's CPU path) on this box's cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PROMPT = [10646, 2501, 263, 931]  # "once upon a time" (reference tokenizer.bin; tests/test_tokenizer.py)
METRIC = "decode_tok_per_s_batch1_f32"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["bf16_tflops"]), "measured"
        except Exception:
            pass
    return 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_mem_available() -> int:
    avail = 0
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                avail = int(ln.split()[1]) * 1024
    except Exception:
        pass
    try:
        mx = open("/sys/fs/cgroup/memory.max").read().strip()
        if mx != "max":
            cur = int(open("/sys/fs/cgroup/memory.current").read().strip())
            avail = min(avail, int(mx) - cur) if avail else int(mx) - cur
    except Exception:
        pass
    return avail


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def oracle_model(cfg, spec):
    """The CPU oracle on ALL host cores: torchrun exports OMP_NUM_THREADS=1 when nproc > 1, which would leave the
    oracle single-threaded (round 1: the reference arm timed out at N = 2/4/8) — set the thread count explicitly.
    Returns (ref module, model or None, cores, reason-if-skipped)."""
    from oracle import ref
    ref.lib().ref_set_threads(host_threads())
    cores = ref.lib().ref_get_threads()
    need = cfg.file_bytes() + (2 << 30)
    have = host_mem_available()
    if have and have < need:
        return ref, None, cores, f"skipped: {need >> 30} GiB host RAM needed, {have >> 30} GiB available"
    return ref, ref.Model(cfg, ref.synth_tensors(cfg, spec)), cores, None


def run_cpu(cfg, spec, n_tokens: int, reps: int, warm: int, want_logits: bool = False, budget_s: float = 0.0):
    """Times the oracle (CPU restatement of engine/src/device/cpu.rs + infer.rs) on all host cores.
    budget_s > 0: the sample (tokens per step) is sized from a 4-token calibration run so that warm + reps steps end
    within about budget_s (never more than n_tokens, never fewer than 4 tokens per step).
    Returns dict(value, cores, tokens, sample, sec_per_rep, logits, min_gap)."""
    ref, om, cores, why = oracle_model(cfg, spec)
    if om is None:
        return {"value": None, "cores": cores, "tokens": None, "sample": why, "sec_per_rep": None, "logits": None, "min_gap": None}
    if budget_s > 0:
        _, _, _, e = ref.generate(om, ref.State(om), PROMPT, 4, 0.0, 0.9)   # also pages the weights in
        _, _, _, e = ref.generate(om, ref.State(om), PROMPT, 4, 0.0, 0.9)
        n_tokens = int(max(4, min(n_tokens, budget_s / max(1, warm + reps) / (e / 4))))
    el, toks, lg, gap = [], None, None, None
    for i in range(warm + reps):
        st = ref.State(om)
        toks, lg, gap, e = ref.generate(om, st, PROMPT, n_tokens, 0.0, 0.9, want_logits=want_logits and i == warm + reps - 1)
        if i >= warm:
            el.append(e)
        del st
    t = sum(el) / len(el)
    return {"value": n_tokens / t, "cores": cores, "tokens": [int(x) for x in toks],
            "sample": f"first {n_tokens} tokens of the {METRIC} workload per step, {reps} timed step(s)",
            "sec_per_rep": t, "logits": lg, "min_gap": None if gap is None else float(gap)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--model", default="llama2-7B")
    ap.add_argument("--tokens", type=int, default=256)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-tokens", type=int, default=0, help="tokens of the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-prefill", action="store_true", help="skip the prompt-prefill (tensor core) measurement")
    ap.add_argument("--no-batched", action="store_true", help="skip the 64-sequence batched-decode measurement")
    ap.add_argument("--no-small", action="store_true", help="skip the stories110M line (second model of BASELINE.json's metric)")
    ap.add_argument("--seed", type=int, default=1234)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from rama_b200 import checkpoint as ck
    cfg = ck.CONFIGS[args.model]
    spec = ck.SynthSpec(seed=args.seed)
    tokens = min(args.tokens, cfg.seq_len)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = (f"{args.model} (dim {cfg.dim}, {cfg.n_layers} layers, {cfg.n_heads} heads, ffn {cfg.hidden_dim}, "
                f"vocab {cfg.vocab_size}) f32 batch-1 greedy decode, {tokens} tokens per step, prompt 'once upon a time'")
    big = cfg.file_bytes() > (4 << 30)
    # both arms print the same `config` (the driver compares them); run-specific facts live under `run`
    config = {"workload": workload, "parallelism": f"tp{world}", "seed": args.seed,
              "l2": "weights streamed per token exceed L2 (no flush needed)" if cfg.weight_bytes_per_token() / world > 200e6
                    else "working set near/below the 126 MB L2: numbers are L2-assisted, reported as is"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        # every step = a bounded sample of the workload (its first n tokens), n sized so that the whole
        # --steps K --warmup W run ends within ~2 min at N = 1 and ~1 min under torchrun (the other ranks' GPUs idle meanwhile)
        r = run_cpu(cfg, spec, args.cpu_tokens or min(tokens, 32 if big else 256), max(args.steps, 1), args.warmup,
                    budget_s=(110.0 if world == 1 else 50.0) if big else 30.0)
        v, t = r["value"], r["sec_per_rep"]
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "tok/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": None if t is None else t * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "run": {"step": f"bounded sample: {r['sample']}",
                        "what": "CPU oracle = C++ restatement of the reference's Rust CPU path "
                                "(engine/src/device/cpu.rs + transformer/infer.rs) on all host cores; the Rust crate "
                                "cannot be built in this image (no cargo/rustc)"},
                "cpu_baseline": {"value": v, "unit": "tok/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": v, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (CUDA)
    import torch
    import torch.distributed as dist
    from rama_b200.engine import GPU, Session

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    tp = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ids = [GPU.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        tp = (rank, world, ids[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gpu = GPU(local_rank, tp)
    gpu.load_synthetic(cfg, spec)
    sess = Session(gpu)

    # warm-up (graph capture happens on the first call)
    for _ in range(args.warmup):
        toks, _ = sess.generate(PROMPT, tokens, 0.0, 0.9)
    # ---- timed: device-resident loop ----
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        toks, ms = sess.generate(PROMPT, tokens, 0.0, 0.9)
        dev_ms += ms
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clk = clocks.stop() if rank == 0 else {}
    # ---- timed: end to end through forward()+sample() with host token feedback ----
    def host_loop():
        token, out = 1, []
        for pos in range(tokens):
            sess.forward(token, pos)
            nxt = PROMPT[pos] if pos < len(PROMPT) else sess.sample(0.0, 0.9)
            out.append(nxt)
            token = nxt
        sess.sync()
        return out
    host_loop()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_toks = host_loop()
    barrier()
    e2e_s = time.perf_counter() - t0

    times = torch.tensor([dev_ms, e2e_s * 1e3, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms = [float(x) for x in times.tolist()]
    total_tokens = args.steps * tokens
    value = total_tokens / (dev_ms * 1e-3)
    e2e_value = total_tokens / (e2e_ms * 1e-3)

    # ---- prompt prefill (BASELINE config 4): 512 prompt rows in one tensor-core pass ----
    pf = None
    if not args.no_prefill:
        n_pf = min(512, cfg.seq_len)
        rng_tokens = [1] + [(7919 * i + 13) % cfg.vocab_size for i in range(1, n_pf)]
        psess = Session(gpu)
        for _ in range(3):
            psess.prefill(rng_tokens, 0)
        barrier()
        pf_ms = [psess.prefill(rng_tokens, 0)[0] for _ in range(5)]
        barrier()
        _, pf_kinds, pf_launches = psess.prefill(rng_tokens, 0, profile=True)
        pt = torch.tensor([sum(pf_ms) / len(pf_ms)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        pf_avg = float(pt.item())
        D, F, L, V = cfg.dim, cfg.hidden_dim, cfg.n_layers, cfg.vocab_size
        gemm_flops = 2.0 * n_pf * L * (4 * D * D + 3 * D * F) / world      # per GPU, f32-equivalent (2MNK)
        attn_flops = 2.0 * 2 * L * D * n_pf * (n_pf + 1) / 2 / world       # causal QK^T + PV
        psess.close()
        pf = {"rows": n_pf, "ms": round(pf_avg, 3), "tok_per_s": round(n_pf / (pf_avg * 1e-3), 1),
              "speedup_vs_per_token_steps": None, "launches": pf_launches,
              "gemm_f32_equiv_tflops_per_gpu": round(gemm_flops / (pf_kinds["gemm"] * 1e-3) / 1e12, 1),
              "gemm_tf32_tflops_issued_per_gpu": round(3 * gemm_flops / (pf_kinds["gemm"] * 1e-3) / 1e12, 1),
              "attn_f32_tflops_per_gpu": round(attn_flops / max(pf_kinds["attn"], 1e-6) / 1e9, 2),
              "ms_by_kind": {k: round(v, 3) for k, v in pf_kinds.items()},
              "what": "rama_prefill: tcgen05 kind::tf32 GEMMs with in-kernel hi/lo split (3 MMAs per k-step, f32-accurate), "
                      "fused RoPE/KV-write and SwiGLU epilogues, causal attention on the warp-level tensor path (mma.sync m16n8k8 "
                      "tf32, the same 3-term split; RAMA_PREFILL_ATTN=cuda: f32 CUDA cores); ms_by_kind from CUDA events "
                      "around every launch"}

    # ---- batched decode (BASELINE config 5): 64 concurrent sequences, one tensor-core pass per step ----
    bd = None
    if not args.no_batched:
        from rama_b200.engine import Batch
        kv_bytes = 2 * cfg.n_layers * cfg.seq_len * cfg.dim * 4 // world + 4 * (cfg.n_heads * cfg.seq_len + 2 * cfg.vocab_size) + (8 << 20)
        free, _ = gpu.mem_info()
        nb = int(max(0, min(64, (free - (10 << 30)) // kv_bytes)))
        if world > 1:  # every rank must run the same batch
            nbt = torch.tensor([nb], dtype=torch.int64, device="cuda")
            dist.all_reduce(nbt, op=dist.ReduceOp.MIN)
            nb = int(nbt.item())
        if nb >= 2:
            bsess = [Session(gpu) for _ in range(nb)]
            batch = Batch(gpu, 64)
            cur = [1] * nb
            b_steps, b_warm = 48, 8
            def bstep(pos):
                batch.forward(bsess, cur, [pos] * nb)
                nxt = batch.sample(bsess, 0.0, 0.9)
                return [PROMPT[pos]] * nb if pos < len(PROMPT) else nxt
            for pos in range(b_warm):
                cur = bstep(pos)
            barrier()
            t0 = time.perf_counter()
            for pos in range(b_warm, b_warm + b_steps):
                cur = bstep(pos)
            barrier()
            b_s = time.perf_counter() - t0
            bt = torch.tensor([b_s], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(bt, op=dist.ReduceOp.MAX)
            b_s = float(bt.item())
            # the same batch with the loop on the device (rama_generate_batch): one graph replay per step, no PCIe traffic
            n_loop = b_warm + b_steps
            batch.generate(bsess, [PROMPT] * nb, n_loop, 0.0, 0.9)
            barrier()
            _, loop_ms = batch.generate(bsess, [PROMPT] * nb, n_loop, 0.0, 0.9)
            lt = torch.tensor([loop_ms], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(lt, op=dist.ReduceOp.MAX)
            loop_ms = float(lt.item())
            step_ms = b_s / b_steps * 1e3
            avg_pos = b_warm + (b_steps - 1) / 2
            wbytes = cfg.weight_bytes_per_token() / world   # per GPU
            kvb = nb * (2 * cfg.n_layers * (avg_pos + 1) * cfg.dim * 4 + 2 * cfg.n_layers * cfg.dim * 4) / world
            bd = {"sequences": nb, "steps": b_steps, "ms_per_step": round(step_ms, 3),
                  "tok_per_s": round(nb / (step_ms * 1e-3), 1), "speedup_vs_batch1": None,
                  "device_loop": {"ms_per_step": round(loop_ms / n_loop, 3), "tok_per_s": round(nb * n_loop / (loop_ms * 1e-3), 1),
                                  "steps": n_loop, "what": "rama_generate_batch: token feedback on the device, CUDA-event time of the "
                                                           "step loop from position 0, max over ranks"},
                  "launches_per_step": batch.launches_per_step(),
                  "hbm_gbs_algorithmic_per_gpu": round((wbytes + kvb) / (step_ms * 1e-3) / 1e9, 1),
                  "tensor_tf32_tflops_issued_per_gpu": round(3 * nb * 2.0 * (wbytes / 4) / (step_ms * 1e-3) / 1e12, 1),
                  "what": "rama_forward_batch + rama_sample_batch per step, host-driven (token ids cross PCIe both ways "
                          "every step), wall clock around the loop; weights stream once per step for all sequences"}
            batch.close()
            for s_ in bsess:
                s_.close()

    # ---- the small models of BASELINE.json's metric (configs[0..1]; 1 GPU): device loop and host-driven e2e ----
    small = {}
    if world == 1 and args.model == "llama2-7B" and not args.no_small:
        hbm_peak, _ = measured_peaks()
        for sname in ("stories110M", "stories15M"):
            scfg = ck.CONFIGS[sname]
            stoks = min(tokens, scfg.seq_len)
            sgpu = GPU(local_rank)
            sgpu.load_synthetic(scfg, spec)
            ss = Session(sgpu)
            for _ in range(args.warmup):
                ss.generate(PROMPT, stoks, 0.0, 0.9)
            torch.cuda.synchronize()
            s_ms = sum(ss.generate(PROMPT, stoks, 0.0, 0.9)[1] for _ in range(args.steps))
            def s_host_loop():
                token = 1
                for pos in range(stoks):
                    ss.forward(token, pos)
                    token = PROMPT[pos] if pos < len(PROMPT) else ss.sample(0.0, 0.9)
                ss.sync()
            s_host_loop()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                s_host_loop()
            s_e2e = time.perf_counter() - t0
            s_val = args.steps * stoks / (s_ms * 1e-3)
            s_bytes = scfg.avg_bytes_per_token(stoks)
            small[sname] = {"model": f"{sname} (dim {scfg.dim}, {scfg.n_layers} layers, {scfg.n_heads} heads, seq {scfg.seq_len})",
                            "value": round(s_val, 1), "unit": "tok/s", "tokens_per_step": stoks,
                            "e2e": round(args.steps * stoks / s_e2e, 1), "us_per_token": round(1e6 / s_val, 1),
                            "launches_per_token": ss.launches_per_step(),
                            "step_frac_of_hbm_roofline": round(s_bytes * s_val / 1e9 / hbm_peak, 4),
                            "note": f"{scfg.weight_bytes_per_token() / 1e6:.0f} MB of weights per token (L2-resident or close): the step "
                                    "is bound by the length of the dependent kernel chain, not by bytes"}
            ss.close(); sgpu.close()

    # ---- per-kernel event timing (un-graphed) at a few positions: dominant-kernel roofline ----
    prof = {}
    for pos in sorted({0, tokens // 4, tokens // 2, 3 * tokens // 4, tokens - 1}):
        for k, (ms, n) in sess.profile_step(int(toks[pos - 1]) if pos else 1, pos).items():
            a = prof.setdefault(k, [0.0, 0])
            a[0] += ms; a[1] += n
    # ---- in-graph timeline (collective under TP): where a token's time goes INSIDE the captured graph — %globaltimer stamps in
    # every kernel of the production step, programmatic dependent launch included (the event pairs above serialise the kernels and
    # add launch latency: they overstate each kernel by ~20 %)
    from rama_b200.engine import step_timeline, summarize_timeline
    tl_pos = tokens // 2
    tl = step_timeline(sess, int(toks[tl_pos - 1]) if tl_pos else 1, tl_pos, 1, 5)
    tl_sum = summarize_timeline(tl)
    tl_step_us = (tl[-1]["end"] - tl[0]["ready"]) * 1e-3

    def teardown():
        # symmetric on every rank: the library's communicator is destroyed collectively
        sess.close(); gpu.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    # ---- parity of THIS run at THIS size against the CPU oracle (outside every timed region) ----
    # N = 1: the oracle decodes the full 256-token workload (≈50 s at 7B on the box's cores; also the cpu_baseline figure):
    #        all tokens must equal the device loop's, logits compared every 16th position on the oracle's own stream.
    # N > 1: rank 0 runs the oracle for 32 tokens, the stream is broadcast, every rank teacher-forces it through
    #        forward() and the gathered logits are compared: the only way a scaling line can carry TP correctness.
    import numpy as np
    cpu = {"value": None, "unit": "tok/s", "cores": None, "kind": "port", "sample": "skipped"}
    tp_parity = None
    if not args.no_cpu:
        n_par = args.cpu_tokens or ((tokens if world == 1 else min(tokens, 32)) if big else min(tokens, 256))
        r = None
        if rank == 0:
            r = run_cpu(cfg, spec, n_par, 1, 0, want_logits=True)
        ctoks = [r["tokens"] if r else None]
        if world > 1:
            dist.broadcast_object_list(ctoks, src=0)
        ctoks = ctoks[0]
        if ctoks is not None:
            every = 16
            worst, token, checked = 0.0, 1, 0
            for pos in range(len(ctoks)):
                sess.forward(token, pos)
                if pos % every == 0 or pos == len(ctoks) - 1:
                    got = sess.logits()   # under TP: all-gather of the vocabulary slices (every rank takes part)
                    if rank == 0:
                        want = r["logits"][pos].astype(np.float64)
                        worst = max(worst, float(np.max(np.abs(got.astype(np.float64) - want)) / max(1.0, float(np.max(np.abs(want))))))
                        checked += 1
                token = ctoks[pos]
            if rank == 0:
                par = {"tokens": len(ctoks), "tokens_match_gpu": ctoks == [int(x) for x in toks[: len(ctoks)]],
                       "logits_max_rel_err_vs_gpu": worst, "logit_positions_checked": checked, "logit_tol": 1e-3,
                       "min_top2_gap": r["min_gap"], "oracle_cores": r["cores"],
                       "what": "greedy tokens of the device-resident loop vs the CPU oracle's, and teacher-forced logits "
                               "(the oracle's own stream through forward()) every 16th position, same weights and prompt"}
                if world == 1:
                    cpu = {"value": r["value"], "unit": "tok/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
                    cpu.update({k: par[k] for k in ("tokens_match_gpu", "logits_max_rel_err_vs_gpu", "logit_positions_checked",
                                                    "logit_tol", "min_top2_gap")})
                else:
                    tp_parity = par
                    cpu["sample"] = "N > 1: see tp_parity (the oracle ran 32 tokens on rank 0 for parity, not as a baseline)"
                    cpu["cores"] = r["cores"]
        elif rank == 0 and r is not None:
            cpu["sample"] = r["sample"]
            cpu["cores"] = r["cores"]

    if rank != 0:
        teardown()
        return

    peak, peak_src = measured_peaks()
    Fl, D = cfg.hidden_dim // world, cfg.dim
    w13_bytes = 4 * (2 * Fl * D + 4 * D + 2 * Fl)  # w1+w3 rows of this rank, x/add/norm-w in, x out, hb/hb2 out
    w13_ms = prof["w13"][0] / max(prof["w13"][1], 1)
    achieved = w13_bytes / (w13_ms * 1e-3) / 1e9
    tot_ms = sum(v[0] for v in prof.values()) or 1.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes/launch from the committed ncu capture
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"w13:{args.model}:tp{world}")
        except Exception:
            traffic = None
    step_bytes = cfg.avg_bytes_per_token(tokens) / world
    roofline = {"bound": "hbm", "kernel": "gemv_fused<ProNorm,RowsW13,EpiSwiGLU> (rmsnorm -> [w1|w3] -> SwiGLU)",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "frac_of_nominal_8TBs": round(achieved / 8000.0, 4),
                "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": w13_bytes,
                "avg_launch_ms": round(w13_ms, 5),
                "timing": "CUDA events around each launch on the session stream, un-graphed (includes launch latency)",
                "kernel_share_of_step": round(prof["w13"][0] / tot_ms, 4),
                "step_achieved": round(step_bytes * value / 1e9, 1),
                "step_frac": round(step_bytes * value / 1e9 / peak, 4),
                "step_frac_of_nominal_8TBs": round(step_bytes * value / 1e9 / 8000.0, 4),
                "step_bytes_per_token_per_gpu": step_bytes}
    kernels = {k: {"ms_per_token": round(v[0] / 5, 4), "launches_per_token": v[1] // 5} for k, v in prof.items() if v[1]}
    w13_tl = tl_sum.get("w13")
    if w13_tl:
        w13_us = w13_tl["chain"] / w13_tl["n"]
        roofline["in_graph"] = {"avg_launch_us": round(w13_us, 2), "achieved": round(w13_bytes / (w13_us * 1e-6) / 1e9, 1),
                                "frac": round(w13_bytes / (w13_us * 1e-6) / 1e9 / peak, 4),
                                "kernel_share_of_step": round(w13_tl["chain"] / tl_step_us, 4),
                                "what": "the same kernel inside the production CUDA graph: its dependency resolved -> the next kernel's "
                                        f"dependency resolved (globaltimer stamps, position {tl_pos})"}
    timeline = {"pos": tl_pos, "step_us": round(tl_step_us, 1),
                "by_kind": {k: {"launches": a["n"], "us_per_token": round(a["chain"], 1), "prologue_us_per_launch": round(a["prologue"] / a["n"], 2),
                                "body_us_per_launch": round(a["body"] / a["n"], 2), "tail_us_per_launch": round(a["tail"] / a["n"], 2)}
                            for k, a in tl_sum.items()},
                "what": "rama_step_timeline: CTA 0 of every kernel of the captured step stamps entry / dependency resolved / prologue done "
                        "(under TP: peer partials arrived) / end; us_per_token = the kernel kind's share of the step's critical path"}

    line = {"metric": METRIC, "value": round(value, 3), "unit": "tok/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "run": {"tp_exchange": None if world == 1 else os.environ.get("RAMA_TP_COMM", "p2p") +
                    (" (one-shot all-reduce over peer memory fused into the wo/w2 GEMV epilogue + next prologue)"
                     if os.environ.get("RAMA_TP_COMM", "p2p") == "p2p" else " (ncclAllReduce in the graph)"),
                    "timing": "CUDA events around each 256-token graph-replay loop, summed over steps, max over ranks",
                    "wall_ms_per_step": round(wall_ms / args.steps, 3),
                    "e2e_matches_device_loop": [int(x) for x in toks] == e2e_toks},
            "clocks": clk,
            "e2e": {"value": round(e2e_value, 3), "unit": "tok/s", "h2d_bytes_per_step": 32 * tokens,
                    "d2h_bytes_per_step": 8 * (tokens - len(PROMPT)),
                    "what": "forward(token,pos)+sample() per token through the C ABI; 32 B ctrl H2D from pinned "
                            "memory and 8 B D2H per token inside the timed region"},
            "gpu_launches": args.steps * tokens * sess.launches_per_step(),
            "roofline": roofline, "kernels": kernels, "timeline": timeline, "cpu_baseline": cpu}
    if pf is not None:
        pf["speedup_vs_per_token_steps"] = round(pf["tok_per_s"] / value, 1)
        bf16, _ = measured_tensor_peak()
        pf["tensor_roofline"] = {"bound": "tensor", "achieved": pf["gemm_tf32_tflops_issued_per_gpu"], "unit": "TFLOP/s",
                                 "peak": round(bf16 / 2, 1), "frac": round(pf["gemm_tf32_tflops_issued_per_gpu"] / (bf16 / 2), 4),
                                 "peak_source": "half of the measured cuBLAS bf16 burst (tf32 runs at half the bf16 rate; "
                                                "MEASURED_PEAKS.json has no tf32 figure)"}
        line["prefill"] = pf
    for sname, sv in small.items():
        line[sname] = sv
    if tp_parity is not None:
        line["tp_parity"] = tp_parity
    if bd is not None:
        bd["speedup_vs_batch1"] = round(bd["tok_per_s"] / value, 1)
        bd["hbm_frac_of_measured_peak"] = round(bd["hbm_gbs_algorithmic_per_gpu"] / peak, 4)
        line["batched_decode"] = bd
    print(json.dumps(line), flush=True)
    teardown()


if __name__ == "__main__":
    main()
