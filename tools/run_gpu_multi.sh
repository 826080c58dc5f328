mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR -m qcnn_gpu_b200.host.multi_gpu --mode frames --qp 27 --frames 10 --height 480 --width 832 --check > gpurun_out/mg_frames_check.log 2>&1; echo rc=$? >> gpurun_out/mg_frames_check.log
timeout 600 $TR -m qcnn_gpu_b200.host.multi_gpu --mode strips --qp 22 --height 1080 --width 1920 --check > gpurun_out/mg_strips_check.log 2>&1; echo rc=$? >> gpurun_out/mg_strips_check.log
timeout 900 $TR -m qcnn_gpu_b200.host.multi_gpu --mode frames --qp 27 --frames 240 --height 2160 --width 3840 --steps 2 > gpurun_out/mg_config4_n$N.log 2>&1; echo rc=$? >> gpurun_out/mg_config4_n$N.log
timeout 900 $TR -m qcnn_gpu_b200.host.multi_gpu --mode strips --qp 22 --height 4320 --width 7680 --steps 5 > gpurun_out/mg_config5_n$N.log 2>&1; echo rc=$? >> gpurun_out/mg_config5_n$N.log
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo rc=$? >> gpurun_out/bench_n$N.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.log 2>&1; echo rc=$? >> gpurun_out/bench_reference.log
for f in mg_frames_check mg_strips_check mg_config4_n$N mg_config5_n$N bench_n$N bench_reference; do echo "== $f"; grep -E "^\{|rc=|Error|error" gpurun_out/$f.log | cut -c1-700 | tail -4; done
