# the whole GPU suite, then the bench table without the CPU and e2e legs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --skip-e2e --skip-cpu > gpurun_out/r2_sb.json 2> gpurun_out/r2_sb.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_sb.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_sb.json').readlines()[-1])
print('C4 value %.4e ms/step %.4f kernel_ms %.4f launches %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches']))
for e in d.get('secondary') or []:
    print('   %-70s value=%.4g ms=%.4g launches/step=%s' % (e['workload'][:70], e.get('value', 0), e.get('ms_per_step', 0), e.get('launches_per_step')))
PY
