# GPU box: does the PLACEMENT of the out-of-line stage functions relative to the kernel's hot code matter?  (libpskd_padN.so: N x ~2.5 KB of
# never-executed instructions between them, built with tools/build_variant.sh padN "-DPSKD_FZ_PAD_CODE=N")
T="python tools/probe/step_trace.py --steps 60 --rounds 1 --no-profiled"
O=gpurun_out/ab_layout.jsonl; : > $O
for rep in 1; do
$T --tag head >> $O 2>> gpurun_out/ab_layout.err
for v in "$@"; do PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_$v.so $T --tag $v >> $O 2>> gpurun_out/ab_layout.err; done
done
python - <<'PY'
import json
for l in open('gpurun_out/ab_layout.jsonl'):
    d = json.loads(l)
    print(d['tag'], 'ms_2nd_half', d['ms_mean_2nd_half'], 'min', d['ms_min'], 'first3', d['ms_first5'][:3], 'MHz', d['nvml']['sm_med'])
PY
tail -3 gpurun_out/ab_layout.err
