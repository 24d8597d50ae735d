#!/usr/bin/env python
"""Runs rama_op_matmul_nt a few times on one shape (for ncu): gemm_one.py M N K variant flags [iters]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rama_b200 import _lib
from rama_b200._lib import check
from rama_b200.engine import GPU, DeviceBuffer
M, N, K, v, fl = [int(x) for x in sys.argv[1:6]]
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
gpu = GPU(0); L = _lib.lib()
a = DeviceBuffer(gpu, M * K); b = DeviceBuffer(gpu, N * K); o = DeviceBuffer(gpu, M * N)
check(L.rama_synth_fill(gpu.h, a.ptr(), M * K, 1, 1, 0, 1.0, 0.0))
check(L.rama_synth_fill(gpu.h, b.ptr(), N * K, 1, 2, 0, 0.02, 0.0))
ms = C.c_float()
check(L.rama_bench_matmul_nt(gpu.h, o.ptr(), a.ptr(), b.ptr(), M, N, K, v, fl, iters, C.byref(ms)))
print("ms", ms.value)
