mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -8
for wl in c3 c4; do
timeout 120 python bench.py --steps 10 --warmup 3 --workload $wl --skip-e2e --skip-cpu > gpurun_out/k_$wl.json 2>gpurun_out/k_$wl.err; tail -c 600 gpurun_out/k_$wl.err; python -c "
import json; d=json.loads(open('gpurun_out/k_$wl.json').readlines()[-1]); print('$wl', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches'])"
done
