mkdir -p gpurun_out
timeout 900 python bench.py --config stills --steps 10 --no-cpu-baseline > gpurun_out/r2_stills_n1.json 2> gpurun_out/r2_stills_n1.err; echo "stills rc=$?"; tail -c 300 gpurun_out/r2_stills_n1.err
timeout 600 python bench.py --config laue --steps 10 --no-cpu-baseline > gpurun_out/r2_laue_n1.json 2> gpurun_out/r2_laue_n1.err; echo "laue rc=$?"; tail -c 300 gpurun_out/r2_laue_n1.err
timeout 600 python bench.py --config dw --steps 10 --no-cpu-baseline > gpurun_out/r2_dw_n1.json 2> gpurun_out/r2_dw_n1.err; echo "dw rc=$?"; tail -c 300 gpurun_out/r2_dw_n1.err
for f in stills laue dw; do python -c "
import json; d=json.load(open('gpurun_out/r2_${f}_n1.json')); print('$f', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['host_prep_s'], d['config']['partition'])"; done
