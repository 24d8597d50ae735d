#!/bin/bash
# Round 2: the pre-split-activation GEMM of the batched step (RAMA_BATCH_PS=1) against the round-1 tile (=0):
# unit tests, the stand-alone GEMM sweep at the 7B decode shapes, and the 64-sequence batched step of bench.py.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_batch.py -x -q 2>&1 | tail -8
timeout 300 python tools/gemm_sweep.py decode > gpurun_out/r2_gemm_sweep_decode.jsonl 2>&1
python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/r2_gemm_sweep_decode.jsonl") if l.startswith("{")]
for r in rows:
    print(r["M"], r["K"], "v", r["variant"], "S", (r["flags"] >> 8) & 255, "ms", r["ms"], "GB/s", r["weight_gbs"])
PY
for ps in 1 0; do
  RAMA_BATCH_PS=$ps timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-prefill --no-small \
      > gpurun_out/r2_bench_batch_ps$ps.json 2> gpurun_out/r2_bench_batch_ps$ps.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_batch_ps$ps.json").read().strip().splitlines()[-1])
print("PS=$ps", d["batched_decode"])
PY
done
