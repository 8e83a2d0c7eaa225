mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
python bench.py --steps 4 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_warp_kernel -s 4 -c 1 -o gpurun_out/prof_step_r1b python bench.py --steps 4 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
