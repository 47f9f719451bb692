# one 8-GPU box: topology, PCIe probes at N=1,2,4,8, multi-GPU tests, the strong-scaling curve of the default workload,
# config5 at N=8 (results under gpurun_out/)
set -x
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
lscpu | head -30 >> gpurun_out/r02_topo.txt; free -g >> gpurun_out/r02_topo.txt; nproc >> gpurun_out/r02_topo.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29500+n)) tools/probe/pcie_rank_probe.py 2>/dev/null | tail -1 >> gpurun_out/r02_pcie_probe.jsonl
done
cat gpurun_out/r02_pcie_probe.jsonl
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -rs 2>&1 | tail -8 > gpurun_out/r2_tests_multi.log; cat gpurun_out/r2_tests_multi.log
timeout 300 psk_soft_b200/lib/demo_box 4096 40000 > gpurun_out/r02_demo_box.txt 2>&1; tail -12 gpurun_out/r02_demo_box.txt
for n in 8 4 2; do
  timeout 600 $TR --nproc-per-node $n --master-port $((29600+n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err
  tail -2 gpurun_out/r02_bench_n$n.err; cut -c1-600 gpurun_out/r02_bench_n$n.json
done
timeout 600 $TR --nproc-per-node 8 --master-port 29700 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --workload config5 > gpurun_out/r02_bench_config5_n8.json 2> gpurun_out/r02_bench_config5_n8.err
cut -c1-600 gpurun_out/r02_bench_config5_n8.json
timeout 600 $TR --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-e2e --scaling weak > gpurun_out/r02_bench_weak_n8.json 2> gpurun_out/r02_bench_weak_n8.err
cut -c1-400 gpurun_out/r02_bench_weak_n8.json
