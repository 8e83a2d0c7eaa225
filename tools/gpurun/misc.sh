mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_drift.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_source_sink.py -m gpu -q -x > gpurun_out/plain_tests.log 2>&1 && timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_source_sink.py tests/test_gpu_strips.py -m gpu -q -x > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/memcheck.log
