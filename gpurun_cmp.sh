mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25
timeout 300 python bench.py --steps 10 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/k2.json 2>gpurun_out/k2.err; tail -c 600 gpurun_out/k2.err; python -c "
import json; d=json.loads(open('gpurun_out/k2.json').readlines()[-1]); print('c3', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches'])"
