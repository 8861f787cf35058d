set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_obs_pp -s 2 -c 1 -o gpurun_out/r2_pp_v1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_pp1.log 2>&1; tail -3 gpurun_out/ncu_pp1.log
CLB_PP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_obs_tc2 -s 2 -c 1 -o gpurun_out/r2_tc2_v2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_tc2b.log 2>&1; tail -3 gpurun_out/ncu_tc2b.log
