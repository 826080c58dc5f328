mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -f -o gpurun_out/prof_fused_v2 $CMD > gpurun_out/ncu_full_run.log 2>&1
tail -3 gpurun_out/ncu_full_run.log; ls -la gpurun_out/*.ncu-rep
