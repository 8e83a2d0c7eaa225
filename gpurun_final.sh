mkdir -p gpurun_out
CMD="timeout 200 python bench.py --steps 6 --warmup 3 --skip-e2e --skip-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1i_c4.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_warp_kernel -s 6 -c 1 -o gpurun_out/prof_step_r1k_c4 $CMD > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
( time timeout 300 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err; tail -c 2600 gpurun_out/bench_default.json
( time timeout 300 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -3 gpurun_out/bench_ref.err; tail -c 1200 gpurun_out/bench_ref.json
