set -x
run() {
  echo "== [$LIBV] :: $*"
  python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline'] or {}
print(round(d['ms_per_step'],3),'ms', round(d['value']/1e3,1),'Gs/s', {k:round(v,3) for k,v in (r.get('kernel_ms_per_step') or {}).items()})
"
}
# the smoke check first: if parity is broken there is no point in timing
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
for v in default old backonly default; do
  if [ "$v" = "default" ]; then unset PSKD_LIB; else export PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_$v.so; fi
  LIBV=$v
  run --workload bank8psk
  run --workload bank8psk --channels 512
  run --workload bank8psk --channels 3072
done
unset PSKD_LIB
run --workload config3
run --workload config2
run --workload config5
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_tests18.log; cat gpurun_out/r2_tests18.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_n1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_r02_fused_18 $CMD > gpurun_out/ncu_n1c.log 2>&1
