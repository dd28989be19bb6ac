# ncu --set full capture of the dominant kernels (1 GPU), raw metric export as text.
# Kernel launch order inside one forward: res0,res1,res2,res3 / up0,up1,up2,up3 -> skip 3 forwards (12) + 1 to land on *1.
mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:resblock_tc_kernel -s 13 -c 1 -o gpurun_out/prof_res1 python bench.py --steps 1 --warmup 3 --cpu-budget 0.5 > gpurun_out/ncu_full_res1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:convt_tc_kernel -s 13 -c 1 -o gpurun_out/prof_up1 python bench.py --steps 1 --warmup 3 --cpu-budget 0.5 > gpurun_out/ncu_full_up1.log 2>&1
for k in res1 up1; do ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_${k}_raw.csv 2>/dev/null; done
ls -la gpurun_out | head -30
