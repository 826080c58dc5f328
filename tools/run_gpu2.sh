mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_fused.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench_fused.log
tail -30 gpurun_out/r2_tests.log; cat gpurun_out/r2_smoke.log; cut -c1-600 gpurun_out/r2_bench_fused.log
