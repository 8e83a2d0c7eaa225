mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c4.json 2>gpurun_out/bench_c4.err; tail -c 800 gpurun_out/bench_c4.err; tail -c 3000 gpurun_out/bench_c4.json
