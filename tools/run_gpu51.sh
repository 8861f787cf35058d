SECONDS=0
timeout 900 python bench.py --config stills --steps 10 --no-cpu-baseline > gpurun_out/r2f_stills_n1.json 2> gpurun_out/r2f_stills_n1.err; echo "stills1 rc=$? wall ${SECONDS}s"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2f_stills_n1.json') if l.startswith('{')][-1]); print('stills_n1', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['host_prep_s'], d.get('row_prep_ms'))"
