mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python scripts/trace_resblock.py > gpurun_out/trace_res.log 2>&1; cat gpurun_out/trace_res.log
