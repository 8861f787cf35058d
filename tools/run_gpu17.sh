# ping-pong kernel (two CTAs per SM, shared dW buffers): parity, then bench against the default
CLB_PP=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -15
CLB_PP=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_pp2.json 2> gpurun_out/r2_bench_pp2.err; tail -c 1500 gpurun_out/r2_bench_pp2.json
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_def.json 2> gpurun_out/r2_bench_def.err; tail -c 600 gpurun_out/r2_bench_def.json
