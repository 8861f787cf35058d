for v in foldonly biasonly; do
CLB_LIB_PATH=tools/libclb_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$v', d['ms_per_step'], d['roofline']['kernel_ms'])"
done
