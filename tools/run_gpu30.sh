for i in 1 2; do timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -80 > gpurun_out/multi_run$i.log; tail -3 gpurun_out/multi_run$i.log; done
grep -n "Error\|error\|raise\|Exception" gpurun_out/multi_run*.log | head -30
