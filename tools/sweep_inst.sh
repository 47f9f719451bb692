# usage (GPU box): bash tools/sweep_inst.sh v1 v2 ...  -> per variant: step time (full bench) and executed warp instructions (reduced run)
for v in "$@"; do
  if [ "$v" = "default" ]; then unset PSKD_LIB; else export PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_$v.so; fi
  t=$(python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2))")
  i=$(ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:k_fused -s 1 -c 1 --csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --samples 128000 2>/dev/null | grep -E "inst_executed|issue_active" | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')
  echo "variant=[$v] ms/step=$t  inst,issue%= $i"
done
unset PSKD_LIB
