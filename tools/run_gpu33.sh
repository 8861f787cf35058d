CLB_LIB_PATH=tools/libclb_biascol.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -12
CLB_LIB_PATH=tools/libclb_biascol.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('biascol', d['ms_per_step'], d['roofline']['kernel_ms'], d['last_metrics'])"
