mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches_run.log 2>&1
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_full_run.log 2>&1
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_final.log 2>&1
ls -la gpurun_out | tail -8; tail -2 gpurun_out/ncu_full_run.log; cut -c1-300 gpurun_out/bench_final.log
