# round 2: ncu capture of step_tile_kernel (bulk-copy staging), C4; plain run first
mkdir -p gpurun_out
cp rmf_crowdsim_b200/_lib/variants/${1:-tile_tma}.so rmf_crowdsim_b200/_lib/librcs.so
timeout 120 python bench.py --steps 6 --warmup 3 --workload c4 --skip-e2e --skip-cpu > gpurun_out/r2_plain_${2:-a}.log 2>&1 || exit 1
tail -c 400 gpurun_out/r2_plain_${2:-a}.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_tile_kernel -s 6 -c 1 -o gpurun_out/prof_tile_r2${2:-a} -f python bench.py --steps 6 --warmup 3 --workload c4 --skip-e2e --skip-cpu > gpurun_out/r2_ncu_${2:-a}.log 2>&1
tail -3 gpurun_out/r2_ncu_${2:-a}.log
