mkdir -p gpurun_out
for rep in 1 2; do for v in A B; do
  QVRCNN_B200_LIB=$PWD/tools/bin/lib$v.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('variant $v rep $rep: %.1f Mpx/s  %.3f ms/step  e2e %.1f  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']))"
done; done > gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
