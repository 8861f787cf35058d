timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "image or il or stills" 2>&1 | tail -2
for bf in 1 0 1 0; do
CLB_BIAS_FEAT=$bf timeout 600 python bench.py --config stills --obs 25000000 --refl 250000 --steps 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('stills share bias_feat=$bf', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3))"
done
