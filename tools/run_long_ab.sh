mkdir -p gpurun_out
for v in $1; do
  QVRCNN_B200_LIB=$PWD/tools/bin/lib_$v.so timeout 200 python bench.py --steps 400 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); s=d['step_ms']; import statistics as st
print('variant $v: first5 %.3f  last100 median %.3f  avg %.3f  clocks %s' % (st.median(s[:5]), st.median(s[-100:]), d['ms_per_step'], d['clocks']))"
done > gpurun_out/long_ab.log 2>&1
cat gpurun_out/long_ab.log
