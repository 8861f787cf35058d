timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --no-cpu-baseline --deterministic | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('det mono', d['ms_per_step'], d['roofline']['kernel_ms'], d['launches_per_step'])"
python bench.py --steps 10 --no-cpu-baseline | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('default mono', d['ms_per_step'], d['roofline']['kernel_ms'], d['launches_per_step'])"
