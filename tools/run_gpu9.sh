timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none -k regex:"k_refl|k_adam|k_dw_prior|k_var_sumsq|k_reduce|k_pack|k_finalize" -s 16 -c 10 -o gpurun_out/r02_small_kernels_dw -f python bench.py --config dw --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log | cut -c1-100
python bench.py --steps 10 --no-cpu-baseline | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['launches_per_step'])"
