CLB_TC3=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -8
CLB_TC3=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/tc3.err | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('tc3', d['ms_per_step'], d['roofline']['kernel_ms'], d['last_metrics'])"; tail -3 gpurun_out/tc3.err
CLB_LIB_PATH=tools/libclb_dwi0.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('dw issuer 0', d['ms_per_step'], d['roofline']['kernel_ms'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('default', d['ms_per_step'], d['roofline']['kernel_ms'], d['last_metrics'])"
