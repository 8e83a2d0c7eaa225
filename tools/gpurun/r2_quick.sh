# GPU suite + C3/C4 lines of the default build (+ optional extra bench args)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=3 > gpurun_out/r2_tests.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_tests.log
for wl in c3 c4; do
f=gpurun_out/r2q_$wl.json
timeout 120 python bench.py --steps 20 --warmup 5 --workload $wl --skip-e2e --skip-cpu --skip-secondary "$@" > $f 2>$f.err; tail -c 300 $f.err; python -c "
import json; d=json.loads(open('$f').readlines()[-1]); print('$wl', '%.4e'%d['value'], round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), d['gpu_launches'])"
done
