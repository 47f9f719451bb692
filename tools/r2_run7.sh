set -x
timeout 1500 python -m pytest tests -x -q -m gpu -rs 2>&1 | tail -12 > gpurun_out/r2_tests7.log
cat gpurun_out/r2_tests7.log
timeout 900 bash tools/r2_sweep7.sh 2>&1 | tee gpurun_out/r2_sweep7.log
