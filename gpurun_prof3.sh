mkdir -p gpurun_out
CMD="python tools/strip_profile.py 8 3"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_strips8.csv $CMD > gpurun_out/ncu4.log 2>&1; tail -2 gpurun_out/plain.log
