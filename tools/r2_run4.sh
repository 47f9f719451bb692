set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_tests4.log
cat gpurun_out/r2_tests4.log
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --channels 512 > gpurun_out/plain512.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fzs_cb -s 6 -c 5 -o gpurun_out/prof_r02_cb512 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --channels 512 > gpurun_out/ncu512.log 2>&1
tail -3 gpurun_out/ncu512.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r02_launches512.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --channels 512 > gpurun_out/ncu512b.log 2>&1
