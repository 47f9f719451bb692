set -x
timeout 1500 python -m pytest tests -x -q -m gpu -rs 2>&1 | tail -25 > gpurun_out/r2_tests5.log
cat gpurun_out/r2_tests5.log
timeout 1500 bash tools/r2_sweep5.sh 2>&1 | tee gpurun_out/r2_sweep5.log
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --channels 512 > gpurun_out/plain512.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fzs_cb -s 6 -c 5 -o gpurun_out/prof_r02_cb512b python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --channels 512 > gpurun_out/ncu512.log 2>&1
tail -3 gpurun_out/ncu512.log
