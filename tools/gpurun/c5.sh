mkdir -p gpurun_out
( time timeout 900 python bench.py --workload c5 --steps 20 --warmup 3 ) > gpurun_out/c5.json 2> gpurun_out/c5.err; tail -5 gpurun_out/c5.err; tail -c 2200 gpurun_out/c5.json
( time timeout 900 python bench.py --workload c5 --c5-zanlungo --steps 20 --warmup 3 --skip-e2e ) > gpurun_out/c5z.json 2> gpurun_out/c5z.err; tail -5 gpurun_out/c5z.err; tail -c 1800 gpurun_out/c5z.json
