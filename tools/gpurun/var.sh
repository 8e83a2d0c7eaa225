mkdir -p gpurun_out
for args in "--workload c2" "--workload c3 --variant lane" "--workload c4 --variant lane" "--workload c2 --kernel 1"; do
python bench.py --steps 20 --warmup 5 $args --skip-cpu > gpurun_out/var.json 2>gpurun_out/var.err; tail -c 300 gpurun_out/var.err; python -c "
import json; d=json.loads(open('gpurun_out/var.json').readlines()[-1]); print('$args', '%.3e'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['mode'], 'e2e %.3e'%d['e2e']['value'], d['config']['nonfinite'], d['config']['oob'])"
done
