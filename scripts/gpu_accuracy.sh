mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python scripts/accuracy_sweep.py > gpurun_out/accuracy_sweep.json 2> gpurun_out/accuracy.err; tail -2 gpurun_out/accuracy.err; python -c "
import json; d=json.load(open('gpurun_out/accuracy_sweep.json')); print(d['summary']); [print(r) for r in d['generator']]"
