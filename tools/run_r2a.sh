# round 2, first GPU call (2 GPUs): the GPU suite incl. the strip tests, multi-GPU checks, bench at N=1 and N=2, the pair-handshake probe
T=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${T}_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${T}_gpus.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_n1.log 2>&1; echo "bench rc=$?" >> gpurun_out/${T}_bench_n1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${T}_bench_n2.log 2>&1; echo "bench2 rc=$?" >> gpurun_out/${T}_bench_n2.log
if [ -x tools/bin/umma_probe2 ]; then timeout 120 tools/bin/umma_probe2 lat > gpurun_out/${T}_probe2_lat.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_probe2_lat.log; fi
tail -3 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log; cut -c1-600 gpurun_out/${T}_bench_n1.log | tail -3; cut -c1-600 gpurun_out/${T}_bench_n2.log | tail -3
