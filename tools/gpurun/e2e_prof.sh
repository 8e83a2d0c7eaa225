mkdir -p gpurun_out
CMD="timeout 200 python bench.py --steps 3 --warmup 3 --skip-cpu"
$CMD > gpurun_out/plain_e.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_e2e.csv $CMD > gpurun_out/ncu_e.log 2>&1
tail -c 200 gpurun_out/plain_e.log; tail -2 gpurun_out/ncu_e.log
