mkdir -p gpurun_out
for e in 0 2 3; do
  echo "== experiment flags $e" 
  QV_FUSED_EXPERIMENT=$e QV_FUSED_PROFILE=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep "fused profile" | head -1
done > gpurun_out/exp.log 2>&1
cat gpurun_out/exp.log
