# the N=1 part of tools/r2_profile_run.sh (after a change that only touches k_fused): smoke, both bench lines, ncu launch list + full capture
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -4 gpurun_out/r02_smoke.log
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err; cut -c1-200 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_n1_steps20.json 2> gpurun_out/r02_bench_n1_steps20.err; cut -c1-200 gpurun_out/r02_bench_n1_steps20.json
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_n1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_n1.csv $CMD > gpurun_out/ncu_n1a.log 2>&1
$CMD > gpurun_out/plain_n1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_r02_fused -f $CMD > gpurun_out/ncu_n1b.log 2>&1
tail -2 gpurun_out/ncu_n1b.log
