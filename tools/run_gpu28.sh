mkdir -p gpurun_out; rm -f gpurun_out/parity_strict_counts.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py --steps 20 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2f_bench_n1.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d.get('host_prep_s'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log | cut -c1-100
for cfg in laue dw; do timeout 600 python bench.py --config $cfg --steps 10 --no-cpu-baseline > gpurun_out/r2f_${cfg}_n1.json 2> gpurun_out/r2f_${cfg}_n1.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2f_${cfg}_n1.json') if l.startswith('{')][-1]); print('$cfg', d['ms_per_step'], d['roofline']['kernel_ms'], d.get('host_prep_s'))"; done
