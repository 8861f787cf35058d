CLB_LIB_PATH=tools/libclb_stsdiv.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "w32 or short_chains or golden or studentt or image_layers or determin" 2>&1 | tail -3
for v in tools/libclb_stsdiv.so careless_b200/libcareless_b200.so; do
CLB_LIB_PATH=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$v', d['ms_per_step'], d['roofline']['kernel_ms'], d['last_metrics'])"
done
