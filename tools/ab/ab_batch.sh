# batched decode: PDL chain on/off, GEMM chunk variants
python -m pytest tests/test_gpu_gemm.py tests/test_gpu_batch.py tests/test_gpu_forward.py -x -q > gpurun_out/pytest_s3c.log 2>&1; tail -3 gpurun_out/pytest_s3c.log
python - <<'PY'
import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
from rama_b200 import _lib
from rama_b200._lib import check
from rama_b200.engine import GPU, DeviceBuffer
gpu = GPU(0); L = _lib.lib()
for (M, N, K, ks) in [(12288, 64, 4096, 3), (4096, 64, 4096, 9), (22016, 64, 4096, 3), (4096, 64, 11008, 9), (32000, 64, 4096, 1)]:
    a = DeviceBuffer(gpu, M * K); b = DeviceBuffer(gpu, N * K); o = DeviceBuffer(gpu, M * N)
    check(L.rama_synth_fill(gpu.h, a.ptr(), M * K, 1, 1, 0, 1.0, 0.0)); check(L.rama_synth_fill(gpu.h, b.ptr(), N * K, 1, 2, 0, 0.02, 0.0))
    for v in (3, 4):
        ms = C.c_float()
        check(L.rama_bench_matmul_nt(gpu.h, o.ptr(), a.ptr(), b.ptr(), M, N, K, v, 2 | (ks << 8), 20, C.byref(ms)))
        print(M, N, K, "ks", ks, "v", v, round(ms.value * 1e3, 1), "us", round(4.0 * M * K / ms.value / 1e6, 0), "GB/s")
    a.free(); b.free(); o.free()
PY
for p in 1 0; do
  RAMA_BATCH_PDL=$p python bench.py --no-cpu --no-prefill --steps 1 --warmup 3 > gpurun_out/bench_batch_pdl$p.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_batch_pdl$p.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("pdl=$p", d["value"], d["batched_decode"])
else:
    print("FAILED"); print(open("gpurun_out/bench_batch_pdl$p.log").read()[-2000:])
PY
done
