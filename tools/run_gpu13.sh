for w in 1 2; do for v in 0 1; do
  CLB_L2_WINDOW=$w CLB_LIB_PATH=$PWD/tools/libclb_exp_b$v.so timeout 200 python bench.py --steps 10 --no-cpu-baseline | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('window $w bias_ones $v', d['ms_per_step'], d['roofline']['kernel_ms'])"
done; done
