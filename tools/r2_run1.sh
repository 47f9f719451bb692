set -x
nvidia-smi -L
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_tests1.log
cat gpurun_out/r2_tests1.log
timeout 900 bash tools/r2_sweep.sh 2>&1 | tee gpurun_out/r2_sweep1.log
