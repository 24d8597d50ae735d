#!/bin/bash
# Round 2: prefill attention on the tensor cores (mma.sync 3xTF32, RAMA_PREFILL_ATTN=mma, default) vs the f32 CUDA-core kernel (=cuda):
# parity tests of every shape, then the per-kind time of a 512-row and a 2048-row prefill at 7B layer shapes (2 layers), at 12 heads
# of 128 (mid-4layer) and at 110M; last, one ncu capture of the default shape.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prefill.py -x -q -s 2>&1 | grep -v "^$" | tail -8
for mode in ${MODES:-cuda mma mma41 mma42 mma44 mma22 mma14}; do
  for spec in "l7-2layer 512" "l7-2layer 2048" "mid-4layer 512" "stories110M 1024"; do
    set -- $spec
    echo "== attn=$mode $1 rows=$2: $(RAMA_PREFILL_ATTN=$mode timeout 300 python tools/prefill_one.py $1 $2 3 profile 2>&1 | tail -1)"
  done
done | tee gpurun_out/r2_attn_mma.txt
if [ "$NCU" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:prefill_attn_mma -c 1 \
    -f -o gpurun_out/r2_attn_mma python tools/prefill_one.py l7-2layer 512 1 > gpurun_out/r2_attn_mma_ncu.log 2>&1
tail -3 gpurun_out/r2_attn_mma_ncu.log
fi
