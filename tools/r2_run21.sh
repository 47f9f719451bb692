set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "task_kernel" 2>&1 | tail -15
run() {
  echo "== [$LIBV] :: $*"
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline'] or {}
print(round(d['ms_per_step'],3),'ms', round(d['value']/1e3,1),'Gs/s', {k:round(v,3) for k,v in (r.get('kernel_ms_per_step') or {}).items()}, d.get('chain'))
"
}
LIBV=uni
run --workload bank8psk --channels 512
export PSKD_FZS_UNI=0; LIBV=two
run --workload bank8psk --channels 512
unset PSKD_FZS_UNI; LIBV=uni
run --workload bank8psk --channels 1024
run --workload config3
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_bank.py -x -q -m gpu -k "uni" 2>&1 | tail -8
