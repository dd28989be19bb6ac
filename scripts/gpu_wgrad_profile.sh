# ncu --set full of the conv_post1 weight-gradient kernel at the scale-0 training shape (first launch of scripts/post1_bwd_time.py)
mkdir -p gpurun_out; cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:post1_wgrad -c 1 -o gpurun_out/r02_wgrad -f python scripts/post1_bwd_time.py > gpurun_out/r02_wgrad_ncu.log 2>&1
ncu -i gpurun_out/r02_wgrad.ncu-rep --page raw --csv > gpurun_out/r02_wgrad_raw.csv 2>/dev/null
python scripts/ncu_table.py gpurun_out/r02_wgrad_raw.csv "ncu --set full, post1_wgrad_tc_kernel at Bt=32, L=128 (scale 0 of a 16 x 8192 training step); B200, round 2" > gpurun_out/r02_wgrad_ncu.txt
python scripts/ncu_key_metrics.py gpurun_out/r02_wgrad_raw.csv >> gpurun_out/r02_wgrad_ncu.txt
python scripts/ncu_hot_sass.py gpurun_out/r02_wgrad.ncu-rep post1_wgrad 25 >> gpurun_out/r02_wgrad_ncu.txt 2>&1
head -40 gpurun_out/r02_wgrad_ncu.txt
