set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s3a.log 2>&1; tail -5 gpurun_out/pytest_s3a.log
for m in stories110M stories15M llama2-7B; do
  for a in cluster split; do
    RAMA_ATTN=$a python bench.py --model $m --no-cpu --no-prefill --no-batched --steps 5 --warmup 3 > gpurun_out/bench_${m}_${a}.log 2>&1
    python - <<PY
import json
l=[x for x in open("gpurun_out/bench_${m}_${a}.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("$m $a", d["value"], d["e2e"]["value"], d["roofline"]["step_frac"], {k:v["ms_per_token"] for k,v in d["kernels"].items()})
else:
    print("$m $a FAILED"); print(open("gpurun_out/bench_${m}_${a}.log").read()[-2000:])
PY
  done
done
