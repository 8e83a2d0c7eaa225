mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for wl in c3 c4 side1024:1 side1024:4 side1024:8; do
f=gpurun_out/w_$(echo $wl | tr ':' '_').json
timeout 120 python bench.py --steps 10 --warmup 3 --workload $wl --skip-e2e --skip-cpu > $f 2>$f.err; tail -c 400 $f.err; python -c "
import json; d=json.loads(open('$f').readlines()[-1]); print('$wl', '%.3e'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches'], d['config'].get('mean_neighbours'))"
done
