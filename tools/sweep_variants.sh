for v in "" bp8 cw6 cw5 ft6 ft4; do
  if [ -z "$v" ]; then unset PSKD_LIB; else export PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_$v.so; fi
  echo "variant=[$v]"
  python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['kernel_ms_per_step'].items()})"
done
unset PSKD_LIB
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 2 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e', d['e2e'])"
