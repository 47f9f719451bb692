# GPU box: the kernels with the back-stage variant inlined (libpskd_spec.so, -DPSKD_FZ_BACK_SPEC) against the default build
PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_spec.so timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
O=gpurun_out/ab_spec.jsonl; : > $O
for rep in 1 2; do
python tools/probe/step_trace.py --steps 500 --rounds 1 --no-profiled --channels 512 --tag head512 >> $O 2>> gpurun_out/ab_spec.err
PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_spec.so python tools/probe/step_trace.py --steps 500 --rounds 1 --no-profiled --channels 512 --tag spec512 >> $O 2>> gpurun_out/ab_spec.err
done
python tools/probe/step_trace.py --steps 60 --rounds 1 --no-profiled --tag head >> $O 2>> gpurun_out/ab_spec.err
PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_spec.so python tools/probe/step_trace.py --steps 60 --rounds 1 --no-profiled --tag spec >> $O 2>> gpurun_out/ab_spec.err
python - <<'PY'
import json
for l in open('gpurun_out/ab_spec.jsonl'):
    d = json.loads(l)
    print(d['tag'], 'ms_2nd_half', d['ms_mean_2nd_half'], 'min', d['ms_min'], 'med', d['ms_med'], 'MHz', d['nvml']['sm_med'])
PY
tail -3 gpurun_out/ab_spec.err
