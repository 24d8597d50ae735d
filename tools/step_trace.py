#!/usr/bin/env python
"""Phase timeline of the persistent decode step kernel: step_trace.py MODEL [pos]
Prints, per phase kind, the mean work time (barrier exit → next barrier arrival, CTA 0) and the mean barrier
time (arrival → exit) in SM cycles and µs at the nominal 1.965 GHz."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rama_b200 import _lib, checkpoint as ck
from rama_b200._lib import check
from rama_b200.engine import GPU, Session

model = sys.argv[1] if len(sys.argv) > 1 else "stories110M"
pos_q = int(sys.argv[2]) if len(sys.argv) > 2 else 100
cfg = ck.CONFIGS[model]
gpu = GPU(0); gpu.load_synthetic(cfg, ck.SynthSpec())
s = Session(gpu)
toks = s.generate([10646, 2501, 263, 931], min(pos_q + 1, cfg.seq_len))[0]
n = 2 * (5 * cfg.n_layers + 1) + 1
buf = (C.c_longlong * n)(); nn = C.c_int32()
acc = {}
REP = 5
for _ in range(REP):
    check(_lib.lib().rama_step_trace(s.h, int(toks[pos_q - 1]), pos_q, buf, n, C.byref(nn)))
    t = np.array(buf[:], dtype=np.int64)
    names = ["qkv", "attn", "wo", "w13", "w2"]
    # stamps: t[0] entry; then for barrier b: t[1+2b] arrival, t[2+2b] exit
    prev = t[0]
    for b in range(5 * cfg.n_layers + 1):
        kind = names[b % 5] if b < 5 * cfg.n_layers else "cls"
        work = t[1 + 2 * b] - prev
        bar = t[2 + 2 * b] - t[1 + 2 * b]
        a = acc.setdefault(kind, [0, 0, 0]); a[0] += work; a[1] += bar; a[2] += 1
        prev = t[2 + 2 * b]
    total = t[-1] - t[0]
ghz = 1.965
print(f"{model} pos {pos_q}: kernel span (CTA 0) {total} cyc = {total / ghz / 1e3:.1f} us")
for k, (w, b, c) in acc.items():
    print(f"  {k:5s} work {w / c:9.0f} cyc ({w / c / ghz / 1e3:6.2f} us)   barrier wait {b / c:8.0f} cyc ({b / c / ghz / 1e3:6.2f} us)   x{c // REP}")
