for wl in side1024:1 side1024:4; do
timeout 200 python bench.py --steps 10 --warmup 3 --workload $wl --skip-e2e --skip-cpu > gpurun_out/wide.json 2>gpurun_out/wide.err; tail -c 300 gpurun_out/wide.err; python -c "
import json; d=json.loads(open('gpurun_out/wide.json').readlines()[-1]); print('$wl', '%.3e'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['candidates_per_agent'], d['config']['mean_neighbours'])"
done
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -2
