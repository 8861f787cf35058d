for v in 0 1; do
  CLB_LIB_PATH=$PWD/tools/libclb_exp_b$v.so timeout 200 python bench.py --steps 10 --no-cpu-baseline | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('bias_ones $v', d['ms_per_step'], d['roofline']['kernel_ms'], d['last_metrics']['loss'])"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
