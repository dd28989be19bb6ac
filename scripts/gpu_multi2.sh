bash scripts/gpu_multi.sh 2
timeout 600 python scripts/latency_configs.py > gpurun_out/latency_configs.json 2> gpurun_out/latency.err; cat gpurun_out/latency_configs.json
