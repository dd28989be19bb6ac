# full GPU check: all GPU tests, smoke, bench, reference arm, launch list
mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'], d['cpu_baseline'])"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
