mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_disc_gpu.py tests/test_generator_gpu.py -m gpu -q -x > gpurun_out/pytest_disc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_disc.log
tail -30 gpurun_out/pytest_disc.log
{ timeout 300 python scripts/msd_time.py; MG_DISC_GROUP=simt timeout 300 python scripts/msd_time.py; } 2>/dev/null > gpurun_out/msd_time.log
cat gpurun_out/msd_time.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/msd_launches.csv python scripts/msd_time.py 16 8192 1 > gpurun_out/msd_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/msd_launches.csv > gpurun_out/msd_launches_summary.txt 2>&1; head -12 gpurun_out/msd_launches_summary.txt
