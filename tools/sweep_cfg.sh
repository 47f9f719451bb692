#!/bin/bash
# usage (GPU box): bash tools/sweep_cfg.sh workload variant...
w=$1; shift
for v in "$@"; do
  if [ "$v" = "default" ]; then unset PSKD_LIB; else export PSKD_LIB=$PWD/psk_soft_b200/lib/libpskd_$v.so; fi
  echo "variant=[$v]"; bash tools/bench_configs.sh $w
done
