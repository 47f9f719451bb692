# final binary on one 8-GPU box: the strong-scaling curve of the default workload, config5 and the multi-GPU tests
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29608 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
tail -2 gpurun_out/r02_bench_n8.err; cut -c1-500 gpurun_out/r02_bench_n8.json
for n in 4 2; do
  timeout 600 $TR --nproc-per-node $n --master-port $((29600+n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err
  tail -2 gpurun_out/r02_bench_n$n.err; cut -c1-400 gpurun_out/r02_bench_n$n.json
done
timeout 600 $TR --nproc-per-node 8 --master-port 29700 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-e2e --workload config5 > gpurun_out/r02_bench_config5_n8.json 2> gpurun_out/r02_bench_config5_n8.err
cut -c1-400 gpurun_out/r02_bench_config5_n8.json
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -rs 2>&1 | tail -4
