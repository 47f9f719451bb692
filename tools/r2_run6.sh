set -x
timeout 1500 python -m pytest tests -x -q -m gpu -rs 2>&1 | tail -25 > gpurun_out/r2_tests6.log
cat gpurun_out/r2_tests6.log
timeout 1500 bash tools/r2_sweep6.sh 2>&1 | tee gpurun_out/r2_sweep6.log
