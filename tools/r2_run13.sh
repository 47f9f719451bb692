set -x
timeout 1500 python -m pytest tests -x -q -m gpu -rs 2>&1 | tail -12 > gpurun_out/r2_tests13.log
cat gpurun_out/r2_tests13.log
timeout 900 bash tools/r2_sweep10.sh 2>&1 | tee gpurun_out/r2_sweep13.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_n1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_r02_fused_sw $CMD > gpurun_out/ncu_n1b.log 2>&1
