mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
export MG_GEN_PATH=tc
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 2 --warmup 3 --cpu-budget 1 > gpurun_out/ncu_bench_tc.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_tc.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:resblock_tc -s 4 -c 4 -o gpurun_out/prof_res_tc python bench.py --steps 1 --warmup 3 --cpu-budget 1 > gpurun_out/ncu_full_tc.log 2>&1
ls -la gpurun_out/
