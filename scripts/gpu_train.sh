mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -x -s > gpurun_out/pytest_train.log 2>&1; grep -n "AssertionError\|assert \|Error\|passed\|failed\|worst" gpurun_out/pytest_train.log | cut -c1-300 | head -30
