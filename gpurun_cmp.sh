mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
for k in 1 2; do python bench.py --steps 10 --warmup 3 --kernel $k --skip-e2e --skip-cpu > gpurun_out/k$k.json 2>gpurun_out/k$k.err; python -c "
import json; d=json.loads(open('gpurun_out/k$k.json').readlines()[-1]); print('kernel $k', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"; done
python bench.py --steps 10 --warmup 3 --variant lane --skip-e2e --skip-cpu > gpurun_out/lane.json; python -c "
import json; d=json.loads(open('gpurun_out/lane.json').readlines()[-1]); print('lane', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
python bench.py --steps 4 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_warp_kernel -s 4 -c 1 -o gpurun_out/prof_step_r1c python bench.py --steps 4 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
