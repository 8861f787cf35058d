timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2f_bench_n1.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['row_prep_ms'], d['roofline']['frac'], d['roofline']['traffic'])"
