# sustained (power-capped) A/B of existing knobs / variants on the bench bank; GPU box: bash tools/probe/sustained_ab.sh
T="python tools/probe/step_trace.py --steps 120 --rounds 2 --no-profiled"
O=gpurun_out/sustained_ab.jsonl; : > $O
$T --tag default >> $O 2>> gpurun_out/sustained_ab.err
PSKD_FZ_CTAS=5 $T --tag ctas5 >> $O 2>> gpurun_out/sustained_ab.err
PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_blocks2.so $T --tag blocks2 >> $O 2>> gpurun_out/sustained_ab.err
PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_blocks2.so PSKD_FZ_CTAS=5 $T --tag blocks2_ctas5 >> $O 2>> gpurun_out/sustained_ab.err
PSKD_FUSED=0 $T --tag staged4096 >> $O 2>> gpurun_out/sustained_ab.err
python tools/probe/step_trace.py --steps 800 --rounds 2 --no-profiled --channels 512 --tag ch512 >> $O 2>> gpurun_out/sustained_ab.err
python tools/probe/step_trace.py --steps 200 --rounds 2 --no-profiled --channels 2048 --tag ch2048 >> $O 2>> gpurun_out/sustained_ab.err
PSKD_FUSED_MIN=1 python tools/probe/step_trace.py --steps 200 --rounds 2 --no-profiled --channels 2048 --tag ch2048_fused >> $O 2>> gpurun_out/sustained_ab.err
python - <<'PY'
import json
for l in open('gpurun_out/sustained_ab.jsonl'):
    d = json.loads(l)
    print(d['tag'], d['round'], 'ms_2nd_half', d['ms_mean_2nd_half'], 'first5', d['ms_first5'][:3], 'J/step', d['joule_per_step'], 'W', d['watt_avg'], d['nvml'])
PY
tail -3 gpurun_out/sustained_ab.err
