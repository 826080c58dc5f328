mkdir -p gpurun_out
for e in ${1:-0}; do
  echo "== experiment flags $e"
  QV_FUSED_EXPERIMENT=$e QV_FUSED_PROFILE=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep -E "fused (profile|trace|stamps)" | head -${2:-9}
done > gpurun_out/prof.log 2>&1
cat gpurun_out/prof.log
