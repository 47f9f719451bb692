python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
T="python tools/probe/step_trace.py --steps 60 --rounds 1 --no-profiled"
O=gpurun_out/ab_last.jsonl; : > $O
for rep in 1 2; do
PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_old.so $T --tag old >> $O 2>> gpurun_out/ab_last.err
$T --tag new >> $O 2>> gpurun_out/ab_last.err
done
python - <<'PY'
import json
for l in open('gpurun_out/ab_last.jsonl'):
    d = json.loads(l)
    print(d['tag'], 'ms_2nd_half', d['ms_mean_2nd_half'], 'min', d['ms_min'], 'first3', d['ms_first5'][:3])
PY
