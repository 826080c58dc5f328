mkdir -p gpurun_out
T=${1:-r6}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${T}_bench.log
QV_FUSED_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "fused (profile|trace|stamps)" | head -17 > gpurun_out/${T}_prof.log
tail -3 gpurun_out/${T}_tests.log; python - <<PY
import json
l=[x for x in open('gpurun_out/${T}_bench.log') if x.startswith('{')]
if l:
    d=json.loads(l[0]); print('value %.1f  ms %.3f  e2e %.1f  frac %.3f clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks']))
else: print(open('gpurun_out/${T}_bench.log').read()[-800:])
PY
cat gpurun_out/${T}_prof.log
