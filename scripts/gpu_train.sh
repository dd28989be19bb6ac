mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python scripts/train_step_profile.py > gpurun_out/train_profile.txt 2> gpurun_out/train_profile.err; echo "rc=$?"; tail -3 gpurun_out/train_profile.err; head -60 gpurun_out/train_profile.txt | cut -c1-230
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_generator_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 30 --warmup 5 --cpu-budget 1 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_q.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'])"
MG_GEN_SLICES=1 timeout 600 python bench.py --steps 30 --warmup 5 --cpu-budget 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('single chain', d['value'], d['ms_per_step'])"
