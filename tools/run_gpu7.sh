mkdir -p gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
run 8 --config stills --steps 10 > gpurun_out/r2_stills_n8.json 2> gpurun_out/r2_stills_n8.err; echo "stills8 rc=$?"; tail -c 400 gpurun_out/r2_stills_n8.err
run 4 --config stills --steps 10 > gpurun_out/r2_stills_n4.json 2> gpurun_out/r2_stills_n4.err; echo "stills4 rc=$?"
run 2 --config stills --steps 10 > gpurun_out/r2_stills_n2.json 2> gpurun_out/r2_stills_n2.err; echo "stills2 rc=$?"
run 8 --steps 20 > gpurun_out/r2_mono_n8.json 2> gpurun_out/r2_mono_n8.err; echo "mono8 rc=$?"; tail -c 300 gpurun_out/r2_mono_n8.err
for f in stills_n8 stills_n4 stills_n2 mono_n8; do python -c "
import json; d=json.load(open('gpurun_out/r2_$f.json')); print('$f', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['host_prep_s'], d['config']['partition']['imbalance'])"; done
