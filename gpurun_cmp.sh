mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
for k in 1 2; do python bench.py --steps 10 --warmup 3 --kernel $k --skip-e2e --skip-cpu > gpurun_out/k$k.json 2>gpurun_out/k$k.err; python -c "
import json; d=json.loads(open('gpurun_out/k$k.json').readlines()[-1]); print('kernel $k', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"; done
python bench.py --steps 10 --warmup 3 --variant lane --skip-e2e --skip-cpu > gpurun_out/lane.json; cut -c1-200 gpurun_out/lane.json
python bench.py --steps 10 --warmup 3 --no-local-plan --skip-e2e --skip-cpu > gpurun_out/nolp.json; cut -c1-200 gpurun_out/nolp.json
python bench.py --steps 5 --warmup 3 --workload c4 --skip-e2e --skip-cpu > gpurun_out/c4.json; cut -c1-200 gpurun_out/c4.json
