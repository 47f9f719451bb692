run() {
  echo "== ${ENVV[*]} :: $*"
  env "${ENVV[@]}" python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline'] or {}
    print(round(d['ms_per_step'],3),'ms', round(d['value']/1e3,1),'Gs/s', {k:round(v,3) for k,v in (r.get('kernel_ms_per_step') or {}).items()}, d['chain'])
except Exception as e:
    print('FAILED', e)
"
}
ENVV=(X=1)
run --workload config5
run --workload bank8psk --channels 3072
run --workload bank8psk --channels 2560
