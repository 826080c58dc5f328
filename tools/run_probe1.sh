mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1_smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r1_bench_layered.log 2>&1; echo "bench rc=$?" >> gpurun_out/r1_bench_layered.log
for t in num thr ts ldst shift; do timeout 120 tools/bin/umma_probe $t > gpurun_out/probe_$t.log 2>&1; echo "rc=$?" >> gpurun_out/probe_$t.log; done
tail -5 gpurun_out/r1_tests.log; cat gpurun_out/r1_smoke.log; tail -3 gpurun_out/probe_num.log
