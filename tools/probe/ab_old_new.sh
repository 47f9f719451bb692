# GPU box: parity suite on the new build, then sustained A/B of libpskd_old.so (previous build) vs the new default
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
T="python tools/probe/step_trace.py --steps 120 --rounds 2 --no-profiled"
O=gpurun_out/ab_old_new.jsonl; : > $O
for rep in 1 2; do
PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_old.so $T --tag old >> $O 2>> gpurun_out/ab_old_new.err
$T --tag new >> $O 2>> gpurun_out/ab_old_new.err
done
PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_old.so python tools/probe/step_trace.py --steps 800 --rounds 2 --no-profiled --channels 512 --tag old512 >> $O 2>> gpurun_out/ab_old_new.err
python tools/probe/step_trace.py --steps 800 --rounds 2 --no-profiled --channels 512 --tag new512 >> $O 2>> gpurun_out/ab_old_new.err
python - <<'PY'
import json
for l in open('gpurun_out/ab_old_new.jsonl'):
    d = json.loads(l)
    print(d['tag'], d['round'], 'ms_2nd_half', d['ms_mean_2nd_half'], 'min', d['ms_min'], 'first3', d['ms_first5'][:3], 'J/step', d['joule_per_step'], 'MHz', d['nvml']['sm_med'])
PY
tail -3 gpurun_out/ab_old_new.err
