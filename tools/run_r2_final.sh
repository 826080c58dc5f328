# round 2 evidence run (1 GPU): tests, smoke, bench (both arms), ncu launch list + full capture of the fused kernel
T=${1:-r2final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/${T}_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/${T}_bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.log 2>&1; echo "ref rc=$?" >> gpurun_out/${T}_bench_reference.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained --no-extra-configs"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/ncu_launches_run.log 2>&1
timeout 300 $CMD > gpurun_out/ncu_plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -f -o gpurun_out/prof_fused_${T} $CMD > gpurun_out/ncu_full_run.log 2>&1
tail -2 gpurun_out/${T}_tests.log; tail -2 gpurun_out/${T}_smoke.log; cut -c1-400 gpurun_out/${T}_bench.log | tail -2; cut -c1-300 gpurun_out/${T}_bench_reference.log | tail -2; ls -la gpurun_out/*.ncu-rep gpurun_out/${T}_launches.csv; tail -3 gpurun_out/ncu_full_run.log
