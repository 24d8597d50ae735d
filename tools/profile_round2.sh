# Round-2 profile on ONE GPU (each ncu pass after its own plain run exited 0):
#   1. the ncu launch list of a short 7B decode (shares of the step per kernel; absolutes are cold-cache and serialised),
#   2. --set full of the dominant GEMV (rmsnorm -> [w1|w3] -> SwiGLU) on 1 GPU,
#   3. the same kernel as ONE RANK OF A TP = 8 STEP runs it (RAMA_TP_SIM=8: 1/8 shards, peer reduction with 8-CTA clusters),
#      which is where roofline.traffic at N = 8 comes from — ncu cannot wrap a multi-rank command.
# usage: profile_round2.sh [list|full]   (default: both)
set -x
mkdir -p gpurun_out
WHAT=${1:-both}
SHORT="python bench.py --tokens 32 --steps 1 --warmup 3 --no-cpu --no-prefill --no-batched --no-small"
if [ "$WHAT" != "full" ]; then
$SHORT > gpurun_out/r2_plain_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/r2_launches_7b.csv $SHORT > gpurun_out/r2_ncu_list.log 2>&1
echo "launch list rc=$?"
fi
if [ "$WHAT" != "list" ]; then
$SHORT > gpurun_out/r2_plain_short2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"RowsW13" -s 100 -c 3 -f -o gpurun_out/r2_prof_w13 $SHORT > gpurun_out/r2_ncu_full_w13.log 2>&1
echo "full capture rc=$?"
export RAMA_TP_SIM=8
$SHORT > gpurun_out/r2_plain_sim8.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"RowsW13|RowsQKV" -s 200 -c 6 -f -o gpurun_out/r2_prof_sim8 $SHORT > gpurun_out/r2_ncu_full_sim8.log 2>&1
echo "sim8 capture rc=$?"
fi
