for pf in 0 2 4 8; do
RAMA_GEMM_PF=$pf python - <<'PY'
import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
from rama_b200 import _lib
from rama_b200._lib import check
from rama_b200.engine import GPU, DeviceBuffer
gpu = GPU(0); L = _lib.lib()
out=[]
for (M, N, K, ks) in [(12288, 64, 4096, 3), (4096, 64, 4096, 9), (22016, 64, 4096, 3), (4096, 64, 11008, 9)]:
    a = DeviceBuffer(gpu, M * K); b = DeviceBuffer(gpu, N * K); o = DeviceBuffer(gpu, M * N)
    check(L.rama_synth_fill(gpu.h, a.ptr(), M * K, 1, 1, 0, 1.0, 0.0)); check(L.rama_synth_fill(gpu.h, b.ptr(), N * K, 1, 2, 0, 0.02, 0.0))
    for v in (3, 2):
        ms = C.c_float()
        check(L.rama_bench_matmul_nt(gpu.h, o.ptr(), a.ptr(), b.ptr(), M, N, K, v, 2 | (ks << 8), 20, C.byref(ms)))
        out.append(f"{M}x{K} v{v} {ms.value*1e3:.1f}us")
    a.free(); b.free(); o.free()
print("pf", os.environ["RAMA_GEMM_PF"], " ".join(out))
PY
done
RAMA_GEMM_PF=4 python tools/gemm_trace.py 12288 64 4096 3 770 2 | tail -8
