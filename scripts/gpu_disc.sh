mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_disc_gpu.py -m gpu -q -x > gpurun_out/pytest_disc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_disc.log
tail -30 gpurun_out/pytest_disc.log
