for i in 1 2 3 4 5 6; do CLB_BENCH_DEBUG=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline 2>gpurun_out/e2e_dbg_$i.err | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('run $i', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))"; grep "\[bench\]" gpurun_out/e2e_dbg_$i.err | cut -c1-150; done
