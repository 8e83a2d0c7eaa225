# the strip / source-sink / graph tests only (single-process transport on one GPU)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_strips.py tests/test_gpu_source_sink.py tests/test_gpu_graphs.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_strip_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_strip_tests.log
