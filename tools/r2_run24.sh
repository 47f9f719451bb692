run() {
  echo "== [$LIBV] :: $*"
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline'] or {}
print(round(d['ms_per_step'],3),'ms', round(d['value']/1e3,1),'Gs/s', {k:round(v,3) for k,v in (r.get('kernel_ms_per_step') or {}).items()})
"
}
for v in default nopf default nopf; do
  if [ "$v" = "default" ]; then unset PSKD_LIB; else export PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_$v.so; fi
  LIBV=$v
  run --workload bank8psk --channels 512
  run --workload config3
  run --workload config2
done
unset PSKD_LIB
run --workload config5
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
