T=${1:-r2d}
mkdir -p gpurun_out
timeout 120 tools/bin/umma_probe3 dual > gpurun_out/${T}_probe3_dual.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_probe3_dual.log
QVRCNN_B200_LIB=$PWD/tools/bin/lib_lprof.so QV_FUSED_PROFILE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sustained --no-extra-configs 2>&1 | grep -E "fused (profile|trace|mma-side)|ms_per_step" | cut -c1-400 | head -30 > gpurun_out/${T}_lprof.log
cat gpurun_out/${T}_probe3_dual.log; cat gpurun_out/${T}_lprof.log
