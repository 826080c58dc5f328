mkdir -p gpurun_out
for e in $1; do
  echo "== experiment flags $e"
  QV_FUSED_EXPERIMENT=$e python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ms_per_step %.3f value %.0f' % (d['ms_per_step'], d['value']))
"
done > gpurun_out/flags.log 2>&1
cat gpurun_out/flags.log
