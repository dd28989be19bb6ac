mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/trace_resblock.py > gpurun_out/trace_res.log 2>&1; grep -v "conv [1-4]" gpurun_out/trace_res.log
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-budget 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
timeout 600 python scripts/compare_stock_pytorch.py > gpurun_out/compare_stock.json 2> gpurun_out/compare_stock.err; cat gpurun_out/compare_stock.json
