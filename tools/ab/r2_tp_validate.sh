#!/bin/bash
# Round 2: tensor parallelism on N GPUs of one box (usage: r2_tp_validate.sh N).  Both launch modes against the oracle
# (torchrun worker, single-process group), the engine CLI with --gpus N on the full llama2-7B, the TP bench line (with
# tp_parity) and the reference arm under torchrun.
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tp.py -x -q -k "[$N]" > gpurun_out/r2_tp${N}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tp${N}_pytest.log
timeout 600 python tools/engine_cli_tp.py --gpus $N --steps 64 > gpurun_out/r2_tp${N}_engine_cli.log 2>&1; echo "engine rc=$?"; tail -6 gpurun_out/r2_tp${N}_engine_cli.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/r2_tp${N}_bench.json 2> gpurun_out/r2_tp${N}_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_tp${N}_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --impl reference --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/r2_tp${N}_ref.json 2> gpurun_out/r2_tp${N}_ref.err; echo "ref rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r2_tp${N}_bench.json", "gpurun_out/r2_tp${N}_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "tp_parity", d.get("tp_parity"), "prefill", (d.get("prefill") or {}).get("ms"),
              (d.get("prefill") or {}).get("ms_by_kind"), "batched", (d.get("batched_decode") or {}).get("ms_per_step"), "cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", e)
PY
