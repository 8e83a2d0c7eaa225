mkdir -p gpurun_out
for k in 1 2; do timeout 600 python tools/strip_overhead.py > gpurun_out/r2_strip_overhead_$k.log 2>&1; echo rc=$?; tail -2 gpurun_out/r2_strip_overhead_$k.log; done
