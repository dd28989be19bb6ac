# ncu --set full over every kernel of one generator forward + one MSD forward (inside the NVTX range "measured").
# MG_GEN_SLICES=1: one chain, so each generator kernel appears once with its whole config-2 grid.
mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
MG_GEN_SLICES=1 timeout 1500 ncu --set full --clock-control none --nvtx --nvtx-include "measured/" -o gpurun_out/prof_all python scripts/one_forward_each.py > gpurun_out/ncu_all.log 2>&1
tail -3 gpurun_out/ncu_all.log
ncu -i gpurun_out/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv 2>/dev/null
python scripts/ncu_key_metrics.py gpurun_out/prof_all_raw.csv > gpurun_out/prof_all_key_metrics.txt; grep -c "^==" gpurun_out/prof_all_key_metrics.txt
python scripts/ncu_table.py gpurun_out/prof_all_raw.csv "ncu --set full, every kernel of ONE generator forward (config 2: B=64, T=32, single chain) followed by ONE multi-scale-discriminator forward (B=16+16, L=8192); B200, round 1" > gpurun_out/ncu_all_kernels.txt
cat gpurun_out/ncu_all_kernels.txt
