run() {
  echo "== [$LIBV] :: $*"
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline'] or {}
print(round(d['ms_per_step'],3),'ms', round(d['value']/1e3,1),'Gs/s', {k:round(v,3) for k,v in (r.get('kernel_ms_per_step') or {}).items()})
"
}
export PSKD_FZS_UNI=1
for lag in 30 60 125 250 500; do
  export PSKD_FZS_UNI_LAG=$lag; LIBV=lag$lag
  run --workload bank8psk --channels 512
done
export PSKD_FZS_UNI_LAG=60
for seg in 1024 4096; do export PSKD_FZS_SEG=$seg; LIBV=seg$seg; run --workload bank8psk --channels 512; done
