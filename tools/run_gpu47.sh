N=$1; shift
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
for cfg in "$@"; do
run $N --config $cfg --steps 10 --no-cpu-baseline > gpurun_out/r2f_${cfg}_n$N.json 2> gpurun_out/r2f_${cfg}_n$N.err; echo "$cfg n$N rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2f_${cfg}_n$N.json') if l.startswith('{')][-1]); print('${cfg}_n$N', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['scaling'])"
done
