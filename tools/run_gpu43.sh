timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for n in 100000 1000000; do timeout 300 python bench.py --obs $n --refl $((n/20)) --steps 200 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('obs $n', 'ms/step', round(d['ms_per_step'],4), 'obs kernel', round(d['roofline']['kernel_ms'],4), 'launches/step', d['launches_per_step'], 'e2e', round(d['e2e']['ms_per_step'],4))"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_small.csv python bench.py --obs 100000 --refl 5000 --steps 3 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_small.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ki].split('(')[0]].append(float(r[vi].replace(',',''))/1e3)
for n,v in sorted(agg.items(), key=lambda x:-sum(x[1])): print(f'{n[:46]:46s} n={len(v):3d} avg {sum(v)/len(v):8.1f} us')
PY
