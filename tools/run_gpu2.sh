set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "w32 or short_chains or frozen_and_eval or golden or laue_studentt or ev11" > gpurun_out/r2_pp_tests.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2_pp_tests.log
timeout 200 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2_bench_pp.json 2> gpurun_out/r2_bench_pp.err; echo "rc=$?"; tail -c 1500 gpurun_out/r2_bench_pp.json; tail -3 gpurun_out/r2_bench_pp.err
CLB_PP=0 timeout 200 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2_bench_tc2b.json 2> gpurun_out/r2_bench_tc2b.err; echo "rc=$?"; tail -c 700 gpurun_out/r2_bench_tc2b.json
