# Round-2 verification after the tensor-core prefill attention: legacy-MMA rate microbenchmark, full GPU suite, smoke(), the bench line.
mkdir -p gpurun_out
timeout 120 tools/ubench/mma_sync_rate > gpurun_out/r2_mma_sync_rate.txt 2>&1; cat gpurun_out/r2_mma_sync_rate.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_final3_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final3_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final3_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_final3_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final3_bench.json 2> gpurun_out/r2_final3_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_final3_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_final3_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "step_frac", "traffic")})
print("prefill", d["prefill"]["ms"], d["prefill"]["ms_by_kind"], d["prefill"]["tensor_roofline"]["frac"])
print("batched", d["batched_decode"]["ms_per_step"], d["batched_decode"]["hbm_frac_of_measured_peak"])
print("small", d["stories110M"]["value"], d["stories15M"]["value"])
PY
