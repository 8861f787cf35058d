timeout 600 python -m pytest tests/test_gpu_prep.py -m gpu -x -q -s 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_prep.py 2>&1 | tail -5
python bench.py --steps 10 --no-cpu-baseline | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('mono', d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d.get('host_prep_s'))"
