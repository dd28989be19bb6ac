mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_tc_gpu.py -m gpu -q > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc.log
tail -4 gpurun_out/pytest_tc.log
timeout 300 python scripts/trace_resblock.py > gpurun_out/trace_res.log 2>&1; cat gpurun_out/trace_res.log
export MG_GEN_PATH=tc
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-budget 2 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_tc.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tc.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
