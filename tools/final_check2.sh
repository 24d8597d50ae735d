# Round-2 end-of-round verification on ONE GPU: full GPU test suite, smoke(), the bench line at the driver's arguments, the reference arm.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_final_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_final_bench.err
timeout 400 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/r2_final_ref.json 2>&1; echo "reference arm rc=$?"; tail -c 400 gpurun_out/r2_final_ref.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_final_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "step_frac", "traffic")}, d["roofline"].get("in_graph"))
print("cpu", d["cpu_baseline"])
print("prefill", d["prefill"]["ms"], d["prefill"]["ms_by_kind"], d["prefill"]["tensor_roofline"]["frac"])
print("batched", d["batched_decode"]["ms_per_step"], d["batched_decode"]["device_loop"], d["batched_decode"]["hbm_frac_of_measured_peak"])
print("small", d["stories110M"]["value"], d["stories15M"]["value"])
PY
