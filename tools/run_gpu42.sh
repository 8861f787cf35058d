run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
SECONDS=0
run 8 --config stills --steps 10 --no-cpu-baseline > gpurun_out/r2f_stills_n8.json 2> gpurun_out/r2f_stills_n8.err; echo "stills8 rc=$? wall ${SECONDS}s"; tail -c 300 gpurun_out/r2f_stills_n8.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2f_stills_n8.json') if l.startswith('{')][-1]); print('stills_n8', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['host_prep_s'], d.get('row_prep_ms'), d['config']['partition']['imbalance'])"
