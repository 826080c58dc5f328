# round 2: skewed schedule (workers release the MMA warp right after their tcgen05.ld's) against the round-1 schedule
T=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_ref_witness.py -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 600 python tools/kernel_ab.py tools/bin/lib_skew0.so tools/bin/lib_skew1.so --steps 30 --reps 3 --sustained 300 > gpurun_out/${T}_kernel_ab.log 2>&1
tail -3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_kernel_ab.log | cut -c1-300
