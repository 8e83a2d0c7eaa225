# gpurun --gpus N -- 'bash tools/gpurun/r2_dist8.sh N': C4 strips (graphs on: phase B replays, NCCL eager), then the C5 stream on strips
n=${1:-8}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
show() { python - <<PY
import json
lines=[l for l in open('$1').read().splitlines() if l.startswith('{')]
d=json.loads(lines[-1])
print('$2 N=%d value %.4e ms/step %.4f kernel_ms %s e2e %s launches %s graph_steps_rank0 %s verified %s' % (d['n_gpus'], d['value'], d['ms_per_step'], (d.get('roofline') or {}).get('kernel_ms'), (d.get('e2e') or {}).get('value'), d['gpu_launches'], d.get('graph_steps_rank0'), d.get('dist_verified')))
PY
}
timeout 300 python bench.py --gpus $n --steps 20 --warmup 5 --skip-e2e > gpurun_out/r2_d8_c4.json 2> gpurun_out/r2_d8_c4.err; echo "rc=$?"; show gpurun_out/r2_d8_c4.json c4
timeout 400 python bench.py --gpus $n --steps 20 --warmup 5 --workload c5 > gpurun_out/r2_d8_c5.json 2> gpurun_out/r2_d8_c5.err; echo "rc=$?"; tail -c 600 gpurun_out/r2_d8_c5.err; show gpurun_out/r2_d8_c5.json c5
