set -x
timeout 1500 python -m pytest tests -x -q -m gpu -rs 2>&1 | tail -12 > gpurun_out/r2_tests12.log
cat gpurun_out/r2_tests12.log
python bench.py > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; tail -2 gpurun_out/r2_bench12.err; cut -c1-200 gpurun_out/r2_bench12.json
python bench.py --impl reference | cut -c1-1500
