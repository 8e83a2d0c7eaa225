# the driver's default bench line (secondary table included) + the reference arm
mkdir -p gpurun_out
( time timeout 900 python bench.py "$@" > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err ) 2>&1 | tail -3; tail -c 600 gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench.json').readlines()[-1])
print('value %.4e  ms/step %.4f  kernel_ms %.4f  e2e %.4e  launches %d graph_steps %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], (d['e2e'] or {}).get('value', 0), d['gpu_launches'], d.get('graph_steps_so_far')))
print('roofline', {k: d['roofline'].get(k) for k in ('achieved','frac','traffic','fp64','step_dram_bytes')})
for e in d.get('secondary') or []:
    print('  %-90s %s' % (e['workload'][:90], ' '.join('%s=%s' % (k, ('%.4g' % v) if isinstance(v, float) else v) for k, v in e.items() if k in ('value','ms_per_step','kernel_ms','frac','graph_steps','wall_s','error','launches_per_step'))))
print('cpu', d['cpu_baseline']['value'], (d.get('cpu_parallel_port') or {}).get('value'))
PY
