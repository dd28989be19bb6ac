mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_disc_gpu.py tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python scripts/train_step_time.py > gpurun_out/train_step.json 2> gpurun_out/train_step.err; echo "rc=$?"; grep "_ms" gpurun_out/train_step.json; tail -2 gpurun_out/train_step.err
timeout 900 python scripts/train_step_profile.py > gpurun_out/train_profile.txt 2> gpurun_out/train_profile.err; echo "rc=$?"; grep "mg::\|Adam\|FunctionBackward  \|convolution_backward  " gpurun_out/train_profile.txt | cut -c1-100,150-250 | head -30
