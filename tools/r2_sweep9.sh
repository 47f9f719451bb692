run() {
  echo "== [$LIBV] ${ENVV[*]} :: $*"
  env "${ENVV[@]}" python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline'] or {}
    print(round(d['ms_per_step'],3),'ms', round(d['value']/1e3,1),'Gs/s', {k:round(v,3) for k,v in (r.get('kernel_ms_per_step') or {}).items()}, d['chain'])
except Exception as e:
    print('FAILED', e)
"
}
ENVV=(X=1)
for v in default f6; do
  if [ "$v" = "default" ]; then unset PSKD_LIB; else export PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_$v.so; fi
  LIBV=$v; run --workload bank8psk
done
for v in default cb5 cb7; do
  if [ "$v" = "default" ]; then unset PSKD_LIB; else export PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_$v.so; fi
  LIBV=$v; run --workload bank8psk --channels 512
done
unset PSKD_LIB; LIBV=default
run --workload config3
run --workload config2
run --workload config1
run --workload config5 --channels 1024
