# last check of the round: smoke(), the whole GPU suite, the reference arm
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_tests.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
