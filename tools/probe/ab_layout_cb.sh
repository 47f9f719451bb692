# GPU box: placement sweep for k_fzs_cb (libpskd_cbpadN.so built with tools/build_variant.sh cbpadN "-DPSKD_FZS_CB_PAD_CODE=N"), 512-channel shard
T="python tools/probe/step_trace.py --steps 400 --rounds 1 --no-profiled --channels 512"
O=gpurun_out/ab_layout_cb.jsonl; : > $O
for rep in 1 2; do
$T --tag head >> $O 2>> gpurun_out/ab_layout_cb.err
for v in "$@"; do PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_$v.so $T --tag $v >> $O 2>> gpurun_out/ab_layout_cb.err; done
done
python - <<'PY'
import json
for l in open('gpurun_out/ab_layout_cb.jsonl'):
    d = json.loads(l)
    print(d['tag'], 'ms_2nd_half', d['ms_mean_2nd_half'], 'min', d['ms_min'], 'med', d['ms_med'], 'MHz', d['nvml']['sm_med'])
PY
tail -3 gpurun_out/ab_layout_cb.err
