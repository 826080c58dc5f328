mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r3_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3_bench_fused.log 2>&1; echo "bench rc=$?" >> gpurun_out/r3_bench_fused.log
QV_FUSED_PROFILE=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r3_prof.log 2>&1
tail -5 gpurun_out/r3_tests.log; cat gpurun_out/r3_smoke.log; cut -c1-400 gpurun_out/r3_bench_fused.log; grep "fused profile" gpurun_out/r3_prof.log | tail -2
