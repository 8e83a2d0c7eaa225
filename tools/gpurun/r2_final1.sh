# round 2, evidence for profiles/: GPU suite (drift rows), launch list, full capture of step_tile_kernel, default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2_tests.log 2>&1; echo "pytest rc=$?"; tail -9 gpurun_out/r2_tests.log
bash tools/gpurun/r2_launches.sh c4
cp rmf_crowdsim_b200/_lib/librcs.so rmf_crowdsim_b200/_lib/variants_final.so 2>/dev/null
timeout 120 python bench.py --steps 6 --warmup 3 --workload c4 --skip-e2e --skip-cpu --skip-secondary > gpurun_out/r2_plain_d.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_tile_kernel -s 6 -c 1 -o gpurun_out/prof_tile_r2d -f python bench.py --steps 6 --warmup 3 --workload c4 --skip-e2e --skip-cpu --skip-secondary > gpurun_out/r2_ncu_d.log 2>&1
tail -2 gpurun_out/r2_ncu_d.log | cut -c1-200
bash tools/gpurun/r2_bench.sh
