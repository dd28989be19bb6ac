mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python scripts/latency_configs.py > gpurun_out/latency_configs.json 2> gpurun_out/latency.err; tail -2 gpurun_out/latency.err; cat gpurun_out/latency_configs.json
