#!/usr/bin/env python
"""Times the tcgen05 3xTF32 GEMM (rama_op_matmul_nt) over the prefill / batched-decode shapes of llama2-7B.
Prints JSON lines: shape, variant, ms, f32-equivalent TFLOP/s (2MNK), tf32 tensor TFLOP/s actually issued (x3)."""
import ctypes as C
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rama_b200 import _lib
from rama_b200._lib import check
from rama_b200.engine import GPU, DeviceBuffer

gpu = GPU(0)
L = _lib.lib()
PREFILL = [(512, 4096, 4096), (512, 11008, 4096), (512, 4096, 11008), (512, 32000, 4096), (2048, 4096, 4096)]
DECODE = [(4096, 64, 4096), (11008, 64, 4096), (4096, 64, 11008), (32000, 64, 4096), (12288, 64, 4096), (22016, 64, 4096)]
MODE = sys.argv[1] if len(sys.argv) > 1 else "all"
SETS = (((PREFILL, [0, 1, 2, 3], 0), (DECODE, [0, 1, 2, 3], 2)) if MODE == "all" else
        ((PREFILL, [1, 4, 5], 0),) if MODE == "prefill" else ((DECODE, [4, 6, 7, 8, 9], 2),))
for shapes, variants, flags in SETS:
    for (M, N, K) in shapes:
        a = DeviceBuffer(gpu, M * K); b = DeviceBuffer(gpu, N * K); o = DeviceBuffer(gpu, M * N)
        check(L.rama_synth_fill(gpu.h, a.ptr(), M * K, 1, 1, 0, 1.0, 0.0))
        check(L.rama_synth_fill(gpu.h, b.ptr(), N * K, 1, 2, 0, 0.02, 0.0))
        for v in variants:
            for fl in ((flags,) if flags else (flags, flags | 1)) + ((flags | (3 << 8), flags | (6 << 8)) if flags else ()):
                ms = C.c_float()
                check(L.rama_bench_matmul_nt(gpu.h, o.ptr(), a.ptr(), b.ptr(), M, N, K, v, fl, 20, C.byref(ms)))
                tf = 2.0 * M * N * K / (ms.value * 1e-3) / 1e12
                print(json.dumps({"M": M, "N": N, "K": K, "variant": v, "flags": fl, "ms": round(ms.value, 4),
                                  "f32_tflops": round(tf, 1), "tf32_tflops_issued": round(3 * tf, 1),
                                  "weight_gbs": round(4.0 * (M if flags else N) * K / (ms.value * 1e-3) / 1e9, 1)}), flush=True)
        a.free(); b.free(); o.free()
