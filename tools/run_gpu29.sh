# 2 GPUs: the multi-GPU parity tests (in-library exchange, product API under WORLD_SIZE=2), then the 2-GPU bench
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
run 2 --steps 20 --no-cpu-baseline > gpurun_out/r2f_mono_n2.json 2> gpurun_out/r2f_mono_n2.err; echo "mono2 rc=$?"; tail -c 300 gpurun_out/r2f_mono_n2.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2f_mono_n2.json') if l.startswith('{')][-1]); print('mono_n2', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['host_prep_s'])"
