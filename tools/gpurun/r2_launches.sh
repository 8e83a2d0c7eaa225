# launch list (per-kernel durations + DRAM bytes) of the default build: bash tools/gpurun/r2_launches.sh c4 [extra bench args]
wl=${1:-c4}; shift
mkdir -p gpurun_out
timeout 200 python bench.py --steps 4 --warmup 3 --workload $wl --skip-e2e --skip-cpu --skip-secondary "$@" > gpurun_out/r2_plain_l_$wl.log 2>&1 || exit 1
tail -c 300 gpurun_out/r2_plain_l_$wl.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_$wl.csv python bench.py --steps 4 --warmup 3 --workload $wl --skip-e2e --skip-cpu --skip-secondary "$@" > gpurun_out/r2_ncu_l_$wl.log 2>&1
tail -2 gpurun_out/r2_ncu_l_$wl.log | cut -c1-300
