# the 1-GPU rows of BASELINE.md section 4 with the current build
mkdir -p gpurun_out
run() { # name, args...
  f=gpurun_out/t_$1.json; shift
  timeout 170 python bench.py --steps 20 --warmup 5 --skip-e2e --skip-cpu "$@" > $f 2>$f.err; tail -c 300 $f.err
  python -c "
import json; d=json.loads(open('$f').readlines()[-1]); r=d['roofline']; c=d['config']
print('$f', '%.3e'%d['value'], 'ms', round(d['ms_per_step'],4), 'kms', r.get('kernel_ms'), 'GB/s', r.get('achieved'), 'frac', r.get('frac'), 'k', c.get('mean_neighbours'), 'fin', c.get('finite_tti_fraction', c.get('frac_finite_tti')), 'nonfinite', c.get('nonfinite'))"
}
run c2 --workload c2
run c3_nolp --workload c3 --no-local-plan
run c3_lane --workload c3 --variant lane
run c3 --workload c3
run c4_nolp --workload c4 --no-local-plan
run c4_lane --workload c4 --variant lane
run c4 --workload c4
run c5 --workload c5 --steps 50
run c5z --workload c5 --c5-zanlungo --steps 50
