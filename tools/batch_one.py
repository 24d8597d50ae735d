#!/usr/bin/env python
"""A few batched decode steps at 7B layer shapes (for ncu launch lists): batch_one.py [model] [B] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rama_b200 import checkpoint as ck
from rama_b200.engine import GPU, Session, Batch
model = sys.argv[1] if len(sys.argv) > 1 else "l7-2layer"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfg = ck.CONFIGS[model]
gpu = GPU(0); gpu.load_synthetic(cfg, ck.SynthSpec())
ss = [Session(gpu) for _ in range(B)]
b = Batch(gpu, 64)
cur = [1] * B
for pos in range(steps):
    b.forward(ss, cur, [pos] * B)
    cur = b.sample(ss, 0.0, 0.9)
print("ok", cur[:4])
