timeout 600 python bench.py --config stills --obs 25000000 --refl 250000 --steps 10 --no-cpu-baseline > gpurun_out/r2f_stills_share.json 2> gpurun_out/r2f_stills_share.err; echo "rc=$?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2f_stills_share.json') if l.startswith('{')][-1]); print('stills share', d['ms_per_step'], d['roofline']['kernel_ms'], d.get('host_prep_s'), d['last_metrics'])"; tail -2 gpurun_out/r2f_stills_share.err
CLB_DEVICE_PREP=0 timeout 600 python bench.py --config stills --obs 25000000 --refl 250000 --steps 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('stills share host prep', d['ms_per_step'], d.get('host_prep_s'), d['last_metrics'])"
