# the round's measurement record on one B200: smoke, full GPU suite, default bench line + reference arm, ncu launch lists and
# full captures of the dominant kernels (each ncu run right after the same command exited 0 without ncu)
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -4 gpurun_out/r02_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu -rs 2>&1 | tail -8 > gpurun_out/r02_tests.log; cat gpurun_out/r02_tests.log
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err; cut -c1-300 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_steps20.json 2> gpurun_out/r02_bench_n1_steps20.err; cut -c1-300 gpurun_out/r02_bench_n1_steps20.json
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-300 gpurun_out/r02_bench_reference.json
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_n1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_n1.csv $CMD > gpurun_out/ncu_n1a.log 2>&1
$CMD > gpurun_out/plain_n1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_r02_fused $CMD > gpurun_out/ncu_n1b.log 2>&1
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --channels 512"
$CMD > gpurun_out/plain_512.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_512.csv $CMD > gpurun_out/ncu_512a.log 2>&1
$CMD > gpurun_out/plain_512.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fzs -s 8 -c 8 -o gpurun_out/prof_r02_fzs512 $CMD > gpurun_out/ncu_512b.log 2>&1
tail -2 gpurun_out/ncu_512b.log
