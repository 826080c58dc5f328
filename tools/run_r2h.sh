# round 2: TMA input ring (QV_FUSED_TMA=1) -- parity, then same-call A/B against the LDG/STS ring
T=${1:-r2h}
mkdir -p gpurun_out
QV_FUSED_TMA=1 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_ref_witness.py -m gpu -x -q > gpurun_out/${T}_tests_tma.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests_tma.log
timeout 600 python tools/kernel_ab.py tools/bin/lib_tma.so tools/bin/lib_tma.so@QV_FUSED_TMA=1 --steps 30 --reps 3 --sustained 300 > gpurun_out/${T}_kernel_ab.log 2>&1
tail -3 gpurun_out/${T}_tests_tma.log; cat gpurun_out/${T}_kernel_ab.log | cut -c1-300
