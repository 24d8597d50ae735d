#!/usr/bin/env python
"""In-graph per-kernel timeline of the decode step (rama_step_timeline): where a token's time goes inside the captured graph.

    python tools/step_timeline.py [--model llama2-7B] [--pos 128] [--reps 8] [--gpus N] [--json out.json]
    RAMA_TP_SIM=8 python tools/step_timeline.py        # one rank's share of a TP = 8 step on ONE GPU (no NVLink, no skew)

--gpus N > 1 uses the single-process multi-device handle (rama_ctx_create_multi); rank 0's stamps are shown.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rama_b200 import checkpoint as ck  # noqa: E402
from rama_b200.engine import GPU, Session, step_timeline, summarize_timeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama2-7B")
    ap.add_argument("--pos", type=int, default=128)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--mode", type=int, default=1)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    cfg = ck.CONFIGS[args.model]
    gpu = GPU.multi(args.gpus) if args.gpus > 1 else GPU(0)
    gpu.load_synthetic(cfg, ck.SynthSpec())
    sess = Session(gpu)
    sess.generate([10646, 2501, 263, 931], min(cfg.seq_len, max(args.pos + 1, 8)), 0.0, 0.9)  # a live cache up to pos, warm graphs
    n_gen = min(cfg.seq_len, 256)
    _, ms = sess.generate([10646, 2501, 263, 931], n_gen, 0.0, 0.9)
    print(f"device-resident loop: {n_gen} tokens in {ms:.2f} ms = {n_gen / ms * 1e3:.1f} tok/s ({ms / n_gen * 1e3:.1f} us per token)")
    tl = step_timeline(sess, 17, args.pos, args.mode, args.reps)
    summ = summarize_timeline(tl)
    step_us = (tl[-1]["end"] - tl[0]["ready"]) * 1e-3
    print(f"{args.model} pos {args.pos}: {len(tl)} kernels, step {step_us:.1f} us inside the graph "
          f"(tp_sim={os.environ.get('RAMA_TP_SIM', '0')}, gpus={args.gpus})")
    print(f"{'kind':8s} {'n':>3s} {'chain us':>9s} {'per launch':>10s} {'prologue':>9s} {'body':>8s} {'tail':>8s} {'early':>8s}")
    for k, a in summ.items():
        n = a["n"]
        print(f"{k:8s} {n:3d} {a['chain']:9.1f} {a['chain'] / n:10.2f} {a['prologue'] / n:9.2f} {a['body'] / n:8.2f} {a['tail'] / n:8.2f} {a['early'] / n:8.2f}")
    if args.json:
        json.dump({"model": args.model, "pos": args.pos, "gpus": args.gpus, "tp_sim": os.environ.get("RAMA_TP_SIM", "0"),
                   "tp_cluster": os.environ.get("RAMA_TP_CLUSTER", "default"), "step_us": step_us, "by_kind": summ,
                   "first_layers": tl[:12]}, open(args.json, "w"), indent=1)
    sess.close(); gpu.close()


if __name__ == "__main__":
    main()
