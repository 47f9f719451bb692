# GPU box: sustained / burst A/B of libpskd_old.so (previous build) vs the new default, no parity suite (run that separately)
T="python tools/probe/step_trace.py --steps 100 --rounds 2 --no-profiled"
O=gpurun_out/ab_quick.jsonl; : > $O
PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_old.so $T --tag old >> $O 2>> gpurun_out/ab_quick.err
$T --tag new >> $O 2>> gpurun_out/ab_quick.err
PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_old.so $T --tag old >> $O 2>> gpurun_out/ab_quick.err
$T --tag new >> $O 2>> gpurun_out/ab_quick.err
PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_old.so python tools/probe/step_trace.py --steps 600 --rounds 2 --no-profiled --channels 512 --tag old512 >> $O 2>> gpurun_out/ab_quick.err
python tools/probe/step_trace.py --steps 600 --rounds 2 --no-profiled --channels 512 --tag new512 >> $O 2>> gpurun_out/ab_quick.err
python - <<'PY'
import json
for l in open('gpurun_out/ab_quick.jsonl'):
    d = json.loads(l)
    print(d['tag'], d['round'], 'ms_2nd_half', d['ms_mean_2nd_half'], 'min', d['ms_min'], 'first3', d['ms_first5'][:3], 'MHz', d['nvml']['sm_med'])
PY
tail -3 gpurun_out/ab_quick.err
