# round 2: linear row chunks for row-window launches + non-unrolled wait loop: parity, strips on 2 GPUs, A/B of the whole-frame kernel
T=${1:-r2k}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 600 python tools/kernel_ab.py tools/bin/lib_tma.so tools/bin/lib_cur.so --steps 30 --reps 3 --sustained 300 > gpurun_out/${T}_kernel_ab.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${T}_bench_n2.log 2>&1; echo "bench rc=$?" >> gpurun_out/${T}_bench_n2.log
tail -3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_kernel_ab.log | cut -c1-300; python - <<'PY'
import json
for l in open('gpurun_out/r2k_bench_n2.log'):
    if l.startswith('{'):
        d=json.loads(l); print('n2 value',round(d['value']),'cfg5',json.dumps(d.get('config5'))[:400])
PY
