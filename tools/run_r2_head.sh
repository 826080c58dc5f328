# HEAD validation (1 GPU): what the driver runs at round end -- GPU tests, smoke, default bench, reference arm
T=${1:-r2head}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
S0=$(date +%s); timeout 900 python bench.py > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench_err.log; echo "bench rc=$? wall=$(( $(date +%s) - S0 ))s" >> gpurun_out/${T}_bench.log
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference.log 2>&1; echo "ref rc=$?" >> gpurun_out/${T}_bench_reference.log
tail -2 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log; cut -c1-300 gpurun_out/${T}_bench.log | tail -2; cut -c1-200 gpurun_out/${T}_bench_reference.log | tail -2
