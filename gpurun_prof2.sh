mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --skip-e2e --skip-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gather_sorted_kernel -s 6 -c 1 -o gpurun_out/prof_gather_r1 $CMD > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
