# ping-pong kernel diagnosis: variants (time only) + one full ncu capture
run() { CLB_PP=1 CLB_LIB_PATH=$1 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$2', d['ms_per_step'], d['roofline']['kernel_ms'])"; }
run careless_b200/libcareless_b200.so pp2
run tools/libclb_pp_nodw.so nodw
run tools/libclb_pp_cta1.so cta1
run tools/libclb_pp_noscr.so noscr
CLB_PP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_obs_pp -s 2 -c 1 -o gpurun_out/r2_pp2_v1 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_pp2.log 2>&1; tail -1 gpurun_out/ncu_pp2.log | cut -c1-100
