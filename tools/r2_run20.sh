set -x
run() {
  echo "== [$LIBV] :: $*"
  python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline'] or {}
print(round(d['ms_per_step'],3),'ms', round(d['value']/1e3,1),'Gs/s', {k:round(v,3) for k,v in (r.get('kernel_ms_per_step') or {}).items()})
"
}
LIBV=auto
run --workload bank8psk
run --workload bank8psk --channels 3072
run --workload bank8psk --channels 3600
export PSKD_FZ_CTAS=5; LIBV=ct5
run --workload bank8psk
run --workload bank8psk --channels 3600
unset PSKD_FZ_CTAS
bash tools/r2_profile_run.sh
