timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --no-cpu-baseline | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('mono', d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"
python bench.py --config stills --obs 25000000 --refl 250000 --steps 10 --no-cpu-baseline | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('stills share', d['ms_per_step'], d['roofline']['kernel_ms'])"
python tools/bench_configs.py --which mono --width 10 --steps 10 | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('mono W10', d['ms_per_step'])"
