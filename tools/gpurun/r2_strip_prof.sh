# per-kernel times of one rank's step at 8 strips (all ranks in one process on one GPU, single-process transport)
mkdir -p gpurun_out
timeout 300 python tools/strip_profile.py 8 3 > gpurun_out/r2_strip_plain.log 2>&1 || { tail -5 gpurun_out/r2_strip_plain.log; exit 1; }
tail -2 gpurun_out/r2_strip_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_strip_launches.csv python tools/strip_profile.py 8 3 > gpurun_out/r2_strip_ncu.log 2>&1
tail -2 gpurun_out/r2_strip_ncu.log
