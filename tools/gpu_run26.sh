mkdir -p gpurun_out
# EE round 0 of config 3 (lane-per-tree kernel), third step: source-level counters
timeout 900 ncu --set full --import-source on --clock-control none -k regex:narrow_round_kernel -s 25 -c 1 -o gpurun_out/r2_ncu_round_c3 -f python tools/launch_list.py c3 > gpurun_out/r2_ncu_round_c3.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2_ncu_round_c3.ncu-rep
