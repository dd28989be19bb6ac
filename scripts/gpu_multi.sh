mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-2}
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_n$N.err
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_n$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref_n$N.json
