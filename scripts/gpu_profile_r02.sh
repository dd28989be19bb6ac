# Round-2 evidence, one GPU: ncu launch list of the bench command, ncu --set full of every kernel of one forward (single chain)
# + one MSD forward, traffic JSON, context timings.  Outputs land in gpurun_out/r02_*; copy the summaries to profiles/.
mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --cpu-budget 0.5 > gpurun_out/r02_ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt; tail -20 gpurun_out/r02_launches_summary.txt
MG_GEN_SLICES=1 timeout 1500 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "measured/" -o gpurun_out/r02_prof_all python scripts/one_forward_each.py > gpurun_out/r02_ncu_all.log 2>&1
tail -2 gpurun_out/r02_ncu_all.log
ncu -i gpurun_out/r02_prof_all.ncu-rep --page raw --csv > gpurun_out/r02_prof_all_raw.csv 2>/dev/null
python scripts/ncu_key_metrics.py gpurun_out/r02_prof_all_raw.csv > gpurun_out/r02_ncu_all_kernels_key_metrics.txt
python scripts/ncu_table.py gpurun_out/r02_prof_all_raw.csv "ncu --set full, every kernel of ONE generator forward (config 2: B=64, T=32, single chain) followed by ONE multi-scale-discriminator forward (B=16+16, L=8192); B200, round 2" > gpurun_out/r02_ncu_all_kernels.txt
cat gpurun_out/r02_ncu_all_kernels.txt
MG_GEN_SLICES=1 python scripts/ncu_traffic_json.py gpurun_out/r02_prof_all_raw.csv > gpurun_out/r02_ncu_traffic.json; head -c 600 gpurun_out/r02_ncu_traffic.json
head -1 gpurun_out/r02_prof_all_raw.csv | tr ',' '\n' | grep -i -E "wavefront|bank_conflict|tensor" > gpurun_out/r02_metric_names.txt
rm -f gpurun_out/r02_prof_all.ncu-rep.tmp
timeout 300 python scripts/trace_resblock.py > gpurun_out/r02_resblock_phase_trace.txt 2>&1
timeout 300 python scripts/latency_configs.py > gpurun_out/r02_latency_configs.json 2>gpurun_out/lat.err
timeout 600 python scripts/train_step_time.py > gpurun_out/r02_train_step_config3.json 2>gpurun_out/train.err
timeout 300 python scripts/msd_time.py > gpurun_out/r02_msd_time.txt 2>&1
timeout 600 python scripts/compare_stock_pytorch.py > gpurun_out/r02_vs_stock_pytorch_gpu.json 2>gpurun_out/cmp.err
ls -la gpurun_out | grep r02
