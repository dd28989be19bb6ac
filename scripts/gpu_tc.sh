mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_tc_gpu.py -m gpu -q > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc.log
tail -40 gpurun_out/pytest_tc.log
MG_GEN_PATH=tc timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget 2 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"
cat gpurun_out/bench_tc.json; tail -3 gpurun_out/bench_tc.err
