mkdir -p gpurun_out
for sch in "" "4,4,8,16,16,8,4,4" "2,6,8,16,16,8,6,2" "2,4,6,10,10,10,10,6,4,2" "8,12,12,12,12,8" "2,2,4,8,16,16,8,4,2,2" "1,3,4,8,16,16,8,4,3,1"; do
  QV_E2E_CHUNKS=$sch timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('chunks [%s]: e2e %.1f  device %.1f' % ('$sch', d['e2e']['value'], d['value']))"
done > gpurun_out/e2e_chunks.log 2>&1
cat gpurun_out/e2e_chunks.log
