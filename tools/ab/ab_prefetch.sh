# A/B of two builds of the library on one box: the working tree's (RoPE factors prefetched per thread) vs librama_b200_base.so (HEAD)
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s3e.log 2>&1; tail -2 gpurun_out/pytest_s3e.log
run() {
  for m in stories110M stories15M; do
    python bench.py --model $m --no-cpu --no-prefill --no-batched --steps 8 --warmup 3 > gpurun_out/tmp_bench.log 2>&1
    python - <<PY
import json
l=[x for x in open("gpurun_out/tmp_bench.log") if x.startswith("{")]
print("$1 $m", json.loads(l[-1])["value"] if l else open("gpurun_out/tmp_bench.log").read()[-800:])
PY
  done
}
run prefetch
cp rama_b200/librama_b200.so /tmp/new.so; cp rama_b200/librama_b200_base.so rama_b200/librama_b200.so
run base
cp /tmp/new.so rama_b200/librama_b200.so
run prefetch
