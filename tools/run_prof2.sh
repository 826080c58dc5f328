mkdir -p gpurun_out
for v in $1; do
  echo "== variant $v"
  QVRCNN_B200_LIB=$PWD/tools/bin/lib_$v.so QV_FUSED_PROFILE=1 timeout 120 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep -E "fused (profile|trace|stamps|mma-side)" | head -28
done > gpurun_out/prof2.log 2>&1
cat gpurun_out/prof2.log
