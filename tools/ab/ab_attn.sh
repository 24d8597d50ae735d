# A/B of the attention variants on one box: RAMA_ATTN_WO = 1 (cluster attention+wo), 2 (per-CTA attention+wo), 0 (separate)
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s3b.log 2>&1; tail -3 gpurun_out/pytest_s3b.log
for m in stories110M stories15M; do
  for a in 1 0; do
    RAMA_ATTN_WO=$a python bench.py --model $m --no-cpu --no-prefill --no-batched --steps 5 --warmup 3 > gpurun_out/bench_${m}_awo${a}.log 2>&1
    python - <<PY
import json
l=[x for x in open("gpurun_out/bench_${m}_awo${a}.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("$m awo=$a", d["value"], d["e2e"]["value"], d["roofline"]["step_frac"], {k:v["ms_per_token"] for k,v in d["kernels"].items()})
else:
    print("$m $a FAILED"); print(open("gpurun_out/bench_${m}_awo${a}.log").read()[-2000:])
PY
  done
done
