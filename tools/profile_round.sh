# Round profile on ONE GPU: the default bench line, then (each after its own plain run exited 0) the ncu launch list of a short
# decode and one --set full capture of the dominant GEMV (w13) and the cluster attention kernel.
set -x
python bench.py > gpurun_out/bench_default_s3.log 2> gpurun_out/bench_default_s3.err; echo "default bench rc=$?"
SHORT="python bench.py --tokens 32 --steps 1 --warmup 3 --no-cpu --no-prefill --no-batched --no-small"
$SHORT > gpurun_out/plain_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/launches_s3.csv $SHORT > gpurun_out/ncu_list_s3.log 2>&1
echo "launch list rc=$?"
$SHORT > gpurun_out/plain_short2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"RowsW13|attn_cluster" -s 200 -c 6 -f -o gpurun_out/prof_s3 $SHORT > gpurun_out/ncu_full_s3.log 2>&1
echo "full capture rc=$?"
S110="python bench.py --model stories110M --tokens 32 --steps 1 --warmup 3 --no-cpu --no-prefill --no-batched"
$S110 > gpurun_out/plain_s110.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 130 --csv --log-file gpurun_out/launches_s110_s3.csv $S110 > gpurun_out/ncu_list_s110.log 2>&1
echo "s110 list rc=$?"
