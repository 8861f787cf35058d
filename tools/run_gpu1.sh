set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc; free -g | head -2
./tools/tc_rate --peak > gpurun_out/tf32_peak.json 2> gpurun_out/tf32_peak.err; cat gpurun_out/tf32_peak.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests1.log 2>&1; tail -15 gpurun_out/r2_gpu_tests1.log
CLB_DISCARD=0 timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2_bench_nodiscard.json 2> gpurun_out/r2_bench_nodiscard.err; tail -c 600 gpurun_out/r2_bench_nodiscard.json
timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2_bench_discard.json 2> gpurun_out/r2_bench_discard.err; tail -c 600 gpurun_out/r2_bench_discard.json
timeout 600 python bench.py --config stills --obs 25000000 --refl 250000 --steps 10 --no-cpu-baseline > gpurun_out/r2_stills_share.json 2> gpurun_out/r2_stills_share.err; tail -c 900 gpurun_out/r2_stills_share.json; tail -5 gpurun_out/r2_stills_share.err
CLB_TC16=0 timeout 600 python bench.py --config stills --obs 25000000 --refl 250000 --steps 5 --no-cpu-baseline > gpurun_out/r2_stills_share_fp32.json 2> gpurun_out/r2_stills_share_fp32.err; tail -c 400 gpurun_out/r2_stills_share_fp32.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_obs_tc2 -s 2 -c 1 -o gpurun_out/r2_tc2_discard -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1; tail -3 gpurun_out/ncu1.log
