# tests, smoke, bench (both arms) for the committed state -- no ncu
T=${1:-f6}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.log 2>&1; echo "ref rc=$?" >> gpurun_out/${T}_bench_reference.log
timeout 600 python bench.py > gpurun_out/${T}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${T}_bench.log
tail -2 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log; cut -c1-300 gpurun_out/${T}_bench.log | tail -2; cut -c1-200 gpurun_out/${T}_bench_reference.log | tail -2
