set -x
nvidia-smi -L
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_tests2.log
cat gpurun_out/r2_tests2.log
timeout 900 bash tools/r2_sweep.sh 2>&1 | tee gpurun_out/r2_sweep2.log
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -3 gpurun_out/r2_bench_default.err; cat gpurun_out/r2_bench_default.json
