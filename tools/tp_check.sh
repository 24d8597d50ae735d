# usage: tp_check.sh N  — multi-GPU parity worker + the 7B TP bench line on N GPUs of one box
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29511 tests/tp_worker.py > gpurun_out/tp${N}_worker_s3.log 2>&1; echo "worker rc=$?"; tail -2 gpurun_out/tp${N}_worker_s3.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_7b_tp${N}_s3.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_7b_tp${N}_s3.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("tp$N", d["value"], d["e2e"]["value"], d["roofline"]["step_frac"], d.get("prefill",{}).get("ms"), d.get("batched_decode",{}).get("ms_per_step"))
else:
    print(open("gpurun_out/bench_7b_tp${N}_s3.log").read()[-3000:])
PY
