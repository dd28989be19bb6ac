mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/utterance_shard_time.py > gpurun_out/utterance_shard_n$N.json 2> gpurun_out/utterance_shard_n$N.err; echo "rc=$?"; tail -3 gpurun_out/utterance_shard_n$N.err | cut -c1-200; cat gpurun_out/utterance_shard_n$N.json
