for kb in 110 208; do
  RAMA_GEMV_STAGE_KB=$kb python bench.py --model mid-4layer --no-cpu --no-prefill --no-batched --steps 5 --warmup 3 > gpurun_out/bench_mid_$kb.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_mid_$kb.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("stage_kb=$kb", d["value"], d["roofline"]["step_frac"], {k:v["ms_per_token"] for k,v in d["kernels"].items()})
else:
    print(open("gpurun_out/bench_mid_$kb.log").read()[-2000:])
PY
done
RAMA_GEMV_STAGE_KB=208 python -m pytest tests/test_gpu_forward.py tests/test_gpu_ops.py -x -q 2>&1 | tail -2
