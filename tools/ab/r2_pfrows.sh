#!/bin/bash
# Round 2: L2 row-burst prefetch of the streamed weight operand (RAMA_GEMM_PFROWS = k-blocks per burst; 0 = round-1 box prefetch)
mkdir -p gpurun_out
for pr in 0 8 16 32; do
  echo "== PFROWS $pr"
  RAMA_GEMM_PFROWS=$pr timeout 300 python tools/gemm_sweep.py decode > gpurun_out/r2_gemm_sweep_pfrows$pr.jsonl 2>&1
  python - <<PY
import json, collections
rows = [json.loads(l) for l in open("gpurun_out/r2_gemm_sweep_pfrows$pr.jsonl") if l.startswith("{")]
best = collections.defaultdict(dict)
for r in rows:
    key = (r["M"], r["K"]); v = r["variant"]
    if v not in best[key] or r["ms"] < best[key][v][0]:
        best[key][v] = (r["ms"], (r["flags"] >> 8) & 255, r["weight_gbs"])
for k, d in best.items():
    print(k, {v: x for v, x in sorted(d.items())})
PY
done
for cfg in "0 0" "8 0" "16 0" "8 1"; do
  set -- $cfg
  RAMA_GEMM_PFROWS=$1 RAMA_BATCH_PS=$2 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-prefill --no-small \
      > gpurun_out/r2_bench_batch_pf$1_ps$2.json 2> gpurun_out/r2_bench_batch_pf$1_ps$2.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_batch_pf$1_ps$2.json").read().strip().splitlines()[-1])
print("PFROWS=$1 PS=$2", d["batched_decode"]["ms_per_step"], d["batched_decode"]["hbm_frac_of_measured_peak"])
PY
done
