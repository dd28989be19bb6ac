# N-GPU checks: NCCL DDP test + bench under torchrun (usage: gpurun --gpus N -- bash scripts/gpu_multi.sh N)
N=${1:-2}
mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -q -s 2>&1 | grep -v Warning | tail -15 | tee gpurun_out/pytest_ddp_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print(d['value'], d['ms_per_step']); print(json.dumps(d.get('multi_gpu'), indent=1))"
