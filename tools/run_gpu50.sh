timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_configs.py tests/test_gpu_model_api.py tests/test_gpu_careless.py -m gpu -x -q 2>&1 | tail -3
for bf in 1 0; do
CLB_BIAS_FEAT=$bf python tools/bench_configs.py --which mono --width 10 --steps 10 | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('mono W10 bias_feat=$bf', d['ms_per_step'])"
CLB_BIAS_FEAT=$bf timeout 600 python bench.py --config stills --obs 25000000 --refl 250000 --steps 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('stills share bias_feat=$bf', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['last_metrics']['loss'], d['last_metrics']['Grad Norm'])"
done
