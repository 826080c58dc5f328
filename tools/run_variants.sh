# usage: tools/run_variants.sh "a b c d" [reps]
mkdir -p gpurun_out
for rep in $(seq 1 ${2:-1}); do for v in $1; do
  QVRCNN_B200_LIB=$PWD/tools/bin/lib_$v.so timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('variant $v rep $rep: %.1f Mpx/s  %.3f ms/step  e2e %.1f  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz']))"
done; done > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
