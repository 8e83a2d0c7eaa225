# round 2, call 1: the whole GPU suite (with the new benchmark-size parity tests) on the round-1 kernels + baseline lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_tests.log
for wl in c3 c4; do
timeout 120 python bench.py --steps 10 --warmup 3 --workload $wl --skip-e2e --skip-cpu > gpurun_out/r2_base_$wl.json 2>gpurun_out/r2_base_$wl.err; tail -c 300 gpurun_out/r2_base_$wl.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_base_$wl.json').readlines()[-1]); print('$wl', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches'])"
done
