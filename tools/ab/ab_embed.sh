python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s3d.log 2>&1; tail -2 gpurun_out/pytest_s3d.log
for m in stories15M stories110M; do
  for e in 0 1; do
    RAMA_EMBED_KERNEL=$e python bench.py --model $m --no-cpu --no-prefill --no-batched --steps 5 --warmup 3 > gpurun_out/bench_${m}_emb$e.log 2>&1
    python - <<PY
import json
l=[x for x in open("gpurun_out/bench_${m}_emb$e.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("$m embed_kernel=$e", d["value"], d["e2e"]["value"], d["roofline"]["step_frac"])
else:
    print(open("gpurun_out/bench_${m}_emb$e.log").read()[-1500:])
PY
  done
done
