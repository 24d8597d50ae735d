mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_prefill.py -x -q 2>&1 | tail -3
for mode in mma mma41; do
  for spec in "l7-2layer 512" "l7-2layer 2048"; do
    set -- $spec
    echo "== attn=$mode $1 rows=$2: $(RAMA_PREFILL_ATTN=$mode timeout 100 python tools/prefill_one.py $1 $2 3 profile 2>&1 | tail -1)"
  done
done | tee gpurun_out/r2_attn_mma_db.txt
