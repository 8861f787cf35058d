timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_configs.py -q -m gpu -x 2>&1 | tail -4
timeout 600 python bench.py --config stills --obs 25000000 --refl 250000 --steps 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('stills share', d['ms_per_step'], d['roofline']['kernel_ms'], d.get('row_prep_ms'))"
python tools/bench_configs.py --which mono --width 10 --steps 10 | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('mono W10', d['ms_per_step'])"
