#!/usr/bin/env python
"""Per-role timeline of one CTA of the tcgen05 3xTF32 GEMM: gemm_trace.py M N K variant flags
Prints, per k-block, clocks relative to the first TMA issue and the mean intervals of the steady state:
  tma   = TMA issue → tile landed (worker sees `full`)        split = landed → split done (`ready` arrive)
  mmaw  = split done → MMA issuer sees `ready`                issue = MMA issuer: ready → 12 MMAs + commit issued
  period= MMA-issue end of k-block kb − that of kb−1          empty = MMAs issued (kb−STAGES) → producer re-issues the stage"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rama_b200 import _lib
from rama_b200._lib import check
from rama_b200.engine import GPU, DeviceBuffer
M, N, K, v, fl = [int(x) for x in sys.argv[1:6]]
stages = int(sys.argv[6]) if len(sys.argv) > 6 else 4
gpu = GPU(0); L = _lib.lib()
a = DeviceBuffer(gpu, M * K); b = DeviceBuffer(gpu, N * K); o = DeviceBuffer(gpu, M * N * 8)
check(L.rama_synth_fill(gpu.h, a.ptr(), M * K, 1, 1, 0, 1.0, 0.0))
check(L.rama_synth_fill(gpu.h, b.ptr(), N * K, 1, 2, 0, 0.02, 0.0))
tr = np.zeros((128, 8), dtype=np.int64)
check(L.rama_debug_gemm_trace(gpu.h, o.ptr(), a.ptr(), b.ptr(), M, N, K, v, fl, tr.ctypes.data_as(C.POINTER(C.c_longlong))))
nkb = int((tr[:, 4] != 0).sum())
t0 = tr[0, 0]
print(f"shape {M}x{N}x{K} variant {v} flags {fl}: {nkb} k-blocks traced (clocks; 1.9 clk/ns)")
print(" kb   tmaIssue  landed  splitDone  mmaReady  mmaIssued | accBack drainWait drainGot")
for kb in range(min(nkb, 40)):
    r = tr[kb]
    f = lambda x: f"{x - t0:8d}" if x else "       -"
    print(f"{kb:3d}  {f(r[0])} {f(r[1])} {f(r[2])}  {f(r[3])}  {f(r[4])}  | {f(r[7])} {f(r[5])} {f(r[6])}")
lo, hi = 8, nkb - 2
if hi > lo + 4:
    s = slice(lo, hi)
    print("steady-state means over k-blocks", lo, "..", hi - 1)
    print("  tma    ", float(np.mean(tr[s, 1] - tr[s, 0])))
    print("  split  ", float(np.mean(tr[s, 2] - tr[s, 1])))
    print("  mmaw   ", float(np.mean(tr[s, 3] - tr[s, 2])))
    print("  issue  ", float(np.mean(tr[s, 4] - tr[s, 3])))
    print("  period ", float(np.mean(tr[lo + 1:hi, 4] - tr[lo:hi - 1, 4])))
    print("  empty  ", float(np.mean(tr[lo + stages:hi, 0] - tr[lo:hi - stages, 4])))
