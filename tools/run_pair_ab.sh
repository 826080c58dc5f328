# parity suite and bench with the CTA-pair kernel (QV_FUSED_PAIR=1) against the default kernel
mkdir -p gpurun_out
QV_FUSED_PAIR=1 timeout 80 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/pair_tests.log 2>&1; echo "pair tests rc=$?"; tail -4 gpurun_out/pair_tests.log
for rep in 1 2; do for pv in 0 1; do
  QV_FUSED_PAIR=$pv timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=sys.stdin.readline()
if l.startswith('{'):
    d=json.loads(l); print('pair=$pv rep $rep: %.1f Mpx/s  %.3f ms/step  e2e %.1f  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz']))
else: print('pair=$pv rep $rep: no result')"
done; done > gpurun_out/pair_ab.log 2>&1
cat gpurun_out/pair_ab.log
