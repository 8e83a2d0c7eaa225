# after a change outside the force kernel: strip / parity tests, launch list, default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_strips.py tests/test_gpu_source_sink.py tests/test_gpu_parity.py tests/test_gpu_graphs.py tests/test_gpu_edges.py -m gpu -x -q > gpurun_out/r2_strip_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_strip_tests.log
bash tools/gpurun/r2_launches.sh c4
bash tools/gpurun/r2_bench.sh
