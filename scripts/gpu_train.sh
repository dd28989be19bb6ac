mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python scripts/train_step_time.py > gpurun_out/train_step.json 2> gpurun_out/train_step.err; echo "rc=$?"; grep "_ms" gpurun_out/train_step.json; tail -2 gpurun_out/train_step.err
