# programmatic dependent launch (RCS_OPT_PDL): the GPU suite with it on, then the bench table with it off / on
mkdir -p gpurun_out
RCS_PDL=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pdl_tests.log 2>&1; echo "pytest (RCS_PDL=1) rc=$?"; tail -4 gpurun_out/r2_pdl_tests.log
for v in 0 1 0 1; do
RCS_PDL=$v timeout 300 python bench.py --steps 20 --warmup 5 --skip-e2e --skip-cpu > gpurun_out/r2_pdl_$v.json 2> gpurun_out/r2_pdl_$v.err; echo "bench RCS_PDL=$v rc=$?"; tail -c 300 gpurun_out/r2_pdl_$v.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_pdl_$v.json').readlines()[-1])
print('PDL=$v C4 value %.4e ms/step %.4f kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))
for e in d.get('secondary') or []:
    print('   %-70s value=%.4g ms=%.4g' % (e['workload'][:70], e.get('value', 0), e.get('ms_per_step', 0)))
PY
done
