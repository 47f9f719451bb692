# usage (on the GPU box): bash tools/sweep_variants.sh v1 v2 ...   ("" = the default build)
for v in "$@"; do
  if [ "$v" = "default" ]; then unset PSKD_LIB; else export PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_$v.so; fi
  echo "variant=[$v]"
  python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['kernel_ms_per_step'].items()}, d['chain'])"
done
unset PSKD_LIB
