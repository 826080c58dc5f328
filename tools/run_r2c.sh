# round 2, third GPU call (1 GPU): GPU suite after the same-device rework, kernel A/B round-1 vs templated row window
T=${1:-r2c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 600 python tools/kernel_ab.py tools/bin/lib_r1.so tools/bin/lib_tmpl.so --steps 30 --reps 3 > gpurun_out/${T}_kernel_ab.log 2>&1
tail -3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_kernel_ab.log | cut -c1-200
