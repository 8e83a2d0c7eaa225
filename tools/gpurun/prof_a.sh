mkdir -p gpurun_out
CMD="timeout 200 python bench.py --steps 6 --warmup 3 --skip-e2e --skip-cpu"
$CMD > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1d_c4.csv $CMD > gpurun_out/ncu_a.log 2>&1
tail -c 300 gpurun_out/plain_a.log; tail -2 gpurun_out/ncu_a.log
