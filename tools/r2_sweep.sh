# usage (on the GPU box): bash tools/r2_sweep.sh   -> one line per configuration (device-resident rate, per-kernel ms)
run() {
  echo "== $*"
  env "${ENVV[@]}" python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline'] or {}
    print(round(d['ms_per_step'],3),'ms', round(d['value']/1e3,1),'Gs/s', 'frac', round((r.get('whole_path') or {}).get('frac',0),3), {k:round(v,3) for k,v in (r.get('kernel_ms_per_step') or {}).items()}, d['chain'], 'launches', d['gpu_launches'])
except Exception as e:
    print('FAILED', e)
"
}
ENVV=(X=1)
run --workload bank8psk
run --workload bank8psk --channels 2048
run --workload bank8psk --channels 1024
run --workload bank8psk --channels 512
run --workload bank8psk --channels 256
run --workload config3
run --workload config2
run --workload config1
ENVV=(PSKD_FUSED=0)
run --workload bank8psk
run --workload bank8psk --channels 2048
ENVV=(PSKD_FUSED=1)
run --workload bank8psk --channels 512
ENVV=(PSKD_FZS=0)
run --workload bank8psk --channels 512
run --workload config3
run --workload config2
