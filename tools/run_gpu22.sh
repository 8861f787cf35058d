CLB_TC3=1 CLB_LIB_PATH=tools/libclb_tc3early.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "w32 or short_chains or golden or studentt" 2>&1 | tail -4
CLB_TC3=1 CLB_LIB_PATH=tools/libclb_tc3early.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/tc3.err | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('tc3 early', d['ms_per_step'], d['roofline']['kernel_ms'], d['last_metrics'])"; tail -3 gpurun_out/tc3.err
