timeout 600 python -m pytest tests/test_protocols.py tests/test_gpu_integration.py -m gpu -x -q 2>&1 | tail -4
for w in 2 0; do
CLB_L2_WINDOW=$w timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_obs_tc2 -s 2 -c 1 --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F, -v w=$w '{print "window",w,$(NF-2),$(NF-1),$NF}'
done
