timeout 600 python -m pytest tests/test_gpu_prep.py -m gpu -x -q -s 2>&1 | tail -6
timeout 900 python tools/bench_prep.py > gpurun_out/r02_prep_times.jsonl 2>gpurun_out/prep.err; cat gpurun_out/r02_prep_times.jsonl; tail -3 gpurun_out/prep.err
python bench.py --steps 10 --no-cpu-baseline | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('mono', d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d.get('host_prep_s'))"
timeout 600 ncu --set full --clock-control none -k regex:"k_prep|k_rs_|k_scan|k_fill" -c 40 -o gpurun_out/r02_prep_kernels -f python tools/bench_prep.py --n 10000000 --child device > gpurun_out/ncu_prep.log 2>&1; tail -2 gpurun_out/ncu_prep.log | cut -c1-150
