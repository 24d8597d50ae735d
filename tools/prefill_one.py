#!/usr/bin/env python
"""One prefill of N rows at 7B layer shapes (for ncu): prefill_one.py [model] [rows] [reps] [profile]
(profile: one more pass with per-kind CUDA-event times)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rama_b200 import checkpoint as ck
from rama_b200.engine import GPU, Session
model = sys.argv[1] if len(sys.argv) > 1 else "l7-2layer"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = ck.CONFIGS[model]
gpu = GPU(0); gpu.load_synthetic(cfg, ck.SynthSpec())
s = Session(gpu)
toks = [1] + [(7919 * i + 13) % cfg.vocab_size for i in range(1, rows)]
for _ in range(reps):
    ms, _, n = s.prefill(toks, 0)
print("ms", ms, "launches", n)
if len(sys.argv) > 4:
    ms, kinds, n = s.prefill(toks, 0, profile=True)
    print("ms", round(ms, 3), "kinds", {k: round(v, 3) for k, v in kinds.items()} if isinstance(kinds, dict) else kinds, "launches", n)
