set -x
export CLB_PP=0
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "w32 or short_chains or frozen_and_eval or golden or laue_studentt or ev11 or image_layers or width32" > gpurun_out/r2_tc2pp_tests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2_tc2pp_tests.log
timeout 200 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2_bench_tc2c.json 2> gpurun_out/r2_bench_tc2c.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_tc2c.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['last_metrics'])"; tail -3 gpurun_out/r2_bench_tc2c.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_obs_tc2 -s 2 -c 1 -o gpurun_out/r2_tc2_v3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_tc2c.log 2>&1; tail -2 gpurun_out/ncu_tc2c.log | cut -c1-200
