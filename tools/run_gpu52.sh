python tools/bench_configs.py --which mono --width 64 --steps 3 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('mono W64', d['ms_per_step'])"
python tools/bench_configs.py --which mono --width 48 --steps 3 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('mono W48', d['ms_per_step'])"
