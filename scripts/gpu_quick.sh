mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_generator_gpu.py -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 5 --cpu-budget 1 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_q.err
python -c "
import json; d=json.load(open('gpurun_out/bench_q.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], {k: round(v*1e3,1) for k,v in d['roofline']['kernel_ms'].items()})"
done
