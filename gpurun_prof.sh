mkdir -p gpurun_out
CMD="timeout 120 python bench.py --steps 4 --warmup 3 --workload c3 --skip-e2e --skip-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1h.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_warp_kernel -s 6 -c 1 -o gpurun_out/prof_step_r1j $CMD > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
