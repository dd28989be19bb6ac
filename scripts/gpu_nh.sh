# A/B of the hand-off split (NH) of the C = 128 / 256 ResBlock kernels: rebuilds the library on the box per variant
mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
for v in "2 1" "4 1" "2 2" "4 2" "4 4"; do set -- $v
  MG_NVCC_EXTRA="-DMG_NH128=$1 -DMG_NH256=$2" python -c "from melgan_multi_b200 import build; build.build(force=True)" > gpurun_out/build_nh.log 2>&1 || { tail -5 gpurun_out/build_nh.log; continue; }
  timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -q -x -k "resblock" 2>&1 | tail -1
  timeout 300 python bench.py --steps 30 --warmup 5 --cpu-budget 1 > gpurun_out/b.json 2>gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('NH128=$1 NH256=$2', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), {k: round(v*1e3,1) for k,v in d['roofline']['kernel_ms'].items()})"
done
