# variants x env knobs on the small-bank workloads (device-resident rate)
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
W512="--workload bank8psk --channels 512"
for v in default cb5 cb8 fr8 fr5; do
  if [ "$v" = "default" ]; then unset PSKD_LIB; else export PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_$v.so; fi
  LIBV=$v
  ENVV=(X=1); run $W512
done
unset PSKD_LIB; LIBV=default
ENVV=(PSKD_FZS_SEG=1024); run $W512
ENVV=(PSKD_FZS_SEG=4096); run $W512
for g in 2 4 8; do
  ENVV=(PSKD_GROUPS=$g); run $W512
  ENVV=(PSKD_GROUPS=$g PSKD_FZS_FRONT_CTAS=3 PSKD_FZS_CB_CTAS=3); run $W512
  ENVV=(PSKD_GROUPS=$g PSKD_FZS_FRONT_CTAS=2 PSKD_FZS_CB_CTAS=4); run $W512
  ENVV=(PSKD_GROUPS=$g PSKD_FZS_FRONT_CTAS=4 PSKD_FZS_CB_CTAS=2); run $W512
done
ENVV=(PSKD_GROUPS=4 PSKD_FZS_FRONT_CTAS=3 PSKD_FZS_CB_CTAS=3); run --workload config3
ENVV=(PSKD_GROUPS=4); run --workload config3
ENVV=(X=1); run --workload config3
ENVV=(PSKD_TP_MAX=8192 PSKD_FUSED=0); run --workload bank8psk
ENVV=(PSKD_TP_MAX=8192 PSKD_FUSED=0 PSKD_GROUPS=8 PSKD_FZS_FRONT_CTAS=3 PSKD_FZS_CB_CTAS=3); run --workload bank8psk
ENVV=(PSKD_TP_MAX=8192 PSKD_FUSED=0 PSKD_GROUPS=8); run --workload bank8psk
ENVV=(X=1); run --workload config5
ENVV=(X=1); run --workload config4
