mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_bench_n$N.log 2>&1; echo rc=$? >> gpurun_out/scale_bench_n$N.log
timeout 900 $TR -m qcnn_gpu_b200.host.multi_gpu --mode frames --qp 27 --frames 240 --height 2160 --width 3840 --steps 3 > gpurun_out/scale_config4_n$N.log 2>&1; echo rc=$? >> gpurun_out/scale_config4_n$N.log
timeout 900 $TR -m qcnn_gpu_b200.host.multi_gpu --mode strips --qp 22 --height 4320 --width 7680 --steps 10 --check > gpurun_out/scale_config5_n$N.log 2>&1; echo rc=$? >> gpurun_out/scale_config5_n$N.log
for f in scale_bench_n$N scale_config4_n$N scale_config5_n$N; do echo "== $f"; grep -E "^\{|rc=" gpurun_out/$f.log | cut -c1-330 | tail -2; done
