timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --no-cpu-baseline | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"
CLB_PP=1 python bench.py --steps 5 --no-cpu-baseline | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('pp', d['ms_per_step'], d['last_metrics']['loss'])"
python bench.py --config laue --steps 5 --no-cpu-baseline | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('laue', d['ms_per_step'])"
