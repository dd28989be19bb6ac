mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_tc_gpu.py -m gpu -q > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc.log
tail -5 gpurun_out/pytest_tc.log
export MG_GEN_PATH=tc
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-budget 2 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tc.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['clocks'], d['cpu_baseline'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 2 --warmup 3 --cpu-budget 1 > gpurun_out/ncu_bench_tc.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_tc.csv
