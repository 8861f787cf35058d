run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
for cfg in laue dw; do
SECONDS=0
run 8 --config $cfg --steps 10 --no-cpu-baseline > gpurun_out/r2f_${cfg}_n8.json 2> gpurun_out/r2f_${cfg}_n8.err; echo "$cfg n8 rc=$? wall ${SECONDS}s"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2f_${cfg}_n8.json') if l.startswith('{')][-1]); print('${cfg}_n8', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['scaling'], d['config']['partition']['imbalance'])"
done
