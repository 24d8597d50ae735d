timeout 120 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -3
timeout 120 python - <<'PY'
import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
from rama_b200 import _lib
from rama_b200._lib import check
from rama_b200.engine import GPU, DeviceBuffer
gpu = GPU(0); L = _lib.lib()
for (M, N, K, ks) in [(12288, 64, 4096, 3), (4096, 64, 4096, 9), (22016, 64, 4096, 3), (4096, 64, 11008, 9), (32000, 64, 4096, 1)]:
    a = DeviceBuffer(gpu, M * K); b = DeviceBuffer(gpu, N * K); o = DeviceBuffer(gpu, M * N)
    check(L.rama_synth_fill(gpu.h, a.ptr(), M * K, 1, 1, 0, 1.0, 0.0)); check(L.rama_synth_fill(gpu.h, b.ptr(), N * K, 1, 2, 0, 0.02, 0.0))
    out=[]
    for v in (3, 5):
        ms = C.c_float()
        check(L.rama_bench_matmul_nt(gpu.h, o.ptr(), a.ptr(), b.ptr(), M, N, K, v, 2 | (ks << 8), 20, C.byref(ms)))
        out.append(f"v{v} {ms.value*1e3:.1f}us {4.0*M*K/ms.value/1e6:.0f}GB/s")
    print(M, K, "ks", ks, " | ".join(out))
    a.free(); b.free(); o.free()
PY
timeout 60 python tools/gemm_trace.py 12288 64 4096 5 770 2 | tail -8
