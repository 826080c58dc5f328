# A/B of library builds on three geometries: the bench workload, one 1080p frame (latency), a tiny frame (per-launch cost)
mkdir -p gpurun_out
L="$@"
timeout 600 python tools/kernel_ab.py $L --steps 30 --reps 3 > gpurun_out/ab3_batch.log 2>&1
timeout 300 python tools/kernel_ab.py $L --steps 200 --reps 2 --frames 1 > gpurun_out/ab3_one_frame.log 2>&1
timeout 300 python tools/kernel_ab.py $L --steps 200 --reps 2 --frames 1 --height 16 --width 120 > gpurun_out/ab3_tiny.log 2>&1
for f in batch one_frame tiny; do echo "== $f"; cut -c1-200 gpurun_out/ab3_$f.log; done
