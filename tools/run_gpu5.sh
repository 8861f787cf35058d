export CLB_PP=0
for v in 0 1 2; do
  CLB_LIB_PATH=$PWD/tools/libclb_exp_v$v.so timeout 200 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r2_order_v$v.json 2> gpurun_out/r2_order_v$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_order_v$v.json')); print('order $v', d['ms_per_step'], d['roofline']['kernel_ms'], d['last_metrics']['loss'])"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "w32 or short_chains or frozen_and_eval or golden or image_layers or width32" 2>&1 | tail -3
