mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
f=gpurun_out/t_c5z.json
timeout 170 python bench.py --steps 50 --warmup 5 --skip-e2e --skip-cpu --workload c5 --c5-zanlungo > $f 2>$f.err; tail -c 300 $f.err
python -c "
import json; d=json.loads(open('$f').readlines()[-1]); r=d['roofline']; c=d['config']
print('$f', '%.3e'%d['value'], 'ms', round(d['ms_per_step'],4), 'kms', r.get('kernel_ms'), 'nonfinite', c.get('nonfinite'), c.get('agents_live'))"
