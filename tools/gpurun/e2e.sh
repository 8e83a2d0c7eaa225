mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 200 python bench.py --skip-cpu > gpurun_out/e2e.json 2> gpurun_out/e2e.err; tail -3 gpurun_out/e2e.err
python -c "
import json; d=json.loads(open('gpurun_out/e2e.json').readlines()[-1]); print('%.4e'%d['value'], d['ms_per_step'], 'e2e %.4e sync %.4e'%(d['e2e']['value'], d['e2e']['synchronous_value']), d['e2e']['buffers_identical'])"
