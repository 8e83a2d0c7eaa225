mkdir -p gpurun_out
CMD="timeout 200 python bench.py --steps 6 --warmup 3 --skip-e2e --skip-cpu"
$CMD > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_warp_kernel -s 6 -c 1 -o gpurun_out/prof_step_r1d_c4 $CMD > gpurun_out/ncu_b.log 2>&1
tail -c 300 gpurun_out/plain_b.log; tail -2 gpurun_out/ncu_b.log
