run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
run 8 --steps 20 > gpurun_out/r2f_mono_n8.json 2> gpurun_out/r2f_mono_n8.err; echo "mono8 rc=$?"; tail -c 300 gpurun_out/r2f_mono_n8.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2f_mono_n8.json') if l.startswith('{')][-1]); print('mono_n8', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['host_prep_s'], d.get('row_prep_ms'))"
run 8 --steps 3 --impl reference > gpurun_out/r2f_ref_n8.json 2> gpurun_out/r2f_ref_n8.err; echo "ref8 rc=$?"; tail -c 400 gpurun_out/r2f_ref_n8.json
