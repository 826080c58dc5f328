# round 2 (N GPUs): multi-GPU strip / shard tests, bench at N
N=${1:-2}; T=${2:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_strips_gpu.py -m gpu -x -q > gpurun_out/${T}_tests_n${N}.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests_n${N}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${T}_bench_n${N}.log 2>&1; echo "bench rc=$?" >> gpurun_out/${T}_bench_n${N}.log
tail -3 gpurun_out/${T}_tests_n${N}.log; cut -c1-300 gpurun_out/${T}_bench_n${N}.log | tail -3
