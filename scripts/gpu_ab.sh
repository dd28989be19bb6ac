# A/B of the generator pipeline knobs (PDL, slices) + parity tests of the tensor-core kernels
mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_generator_gpu.py -m gpu -q -x 2>&1 | tail -2
for p in 0 1; do for sl in 1 4; do MG_PDL=$p MG_GEN_SLICES=$sl timeout 300 python bench.py --steps 30 --warmup 5 --cpu-budget 1 > gpurun_out/b.json 2>gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('PDL=$p slices=$sl', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"; done; done
for p in 0 1; do echo "latency PDL=$p"; MG_PDL=$p timeout 300 python scripts/latency_configs.py 2>&1 | tail -12; done
