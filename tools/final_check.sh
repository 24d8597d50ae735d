# Round-end verification on ONE GPU: full GPU test suite, smoke(), the default bench line, the reference arm, and (after a plain
# run of the same short command) one --set full capture of the dominant GEMV.
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final_s3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final_s3.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_s3.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_s3.log
python bench.py > gpurun_out/bench_default_final_s3.log 2> gpurun_out/bench_default_final_s3.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_s3.log 2>&1; echo "reference arm rc=$?"; tail -c 600 gpurun_out/bench_reference_s3.log
SHORT="python bench.py --tokens 32 --steps 1 --warmup 3 --no-cpu --no-prefill --no-batched --no-small"
$SHORT > gpurun_out/plain_short3.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:RowsW13 -s 100 -c 3 -f -o gpurun_out/prof_w13_s3 $SHORT > gpurun_out/ncu_w13_s3.log 2>&1
echo "w13 capture rc=$?"
