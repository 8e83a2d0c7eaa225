mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
for wl in c3 c4; do
python bench.py --steps 20 --warmup 5 --workload $wl --no-local-plan --skip-e2e --skip-cpu > gpurun_out/nolp_$wl.json 2>gpurun_out/nolp_$wl.err; tail -c 300 gpurun_out/nolp_$wl.err; python -c "
import json; d=json.loads(open('gpurun_out/nolp_$wl.json').readlines()[-1]); print('nolp $wl', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['gpu_launches'])"
done
