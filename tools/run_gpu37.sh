SECONDS=0; timeout 600 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench wall ${SECONDS}s"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2f_bench_n1.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['cpu_baseline']['sample'][:60], d.get('row_prep_ms'))"
SECONDS=0; timeout 600 python bench.py --impl reference > gpurun_out/r2f_ref_n1.json 2> gpurun_out/r2f_ref_n1.err; echo "ref wall ${SECONDS}s"; tail -c 500 gpurun_out/r2f_ref_n1.json
