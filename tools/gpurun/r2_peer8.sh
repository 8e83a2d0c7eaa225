# N-GPU bench line with the peer-store transport, e2e host-time split printed (RCS_E2E_TRACE)
n=${1:-8}; shift
mkdir -p gpurun_out
export NCCL_DEBUG=WARN RCS_E2E_TRACE=1
timeout 600 python bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/r2_peer8_$n.json 2> gpurun_out/r2_peer8_$n.err; echo "bench rc=$?"
grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/r2_peer8_$n.err | tail -c 1500
python - <<PY
import json
lines=[l for l in open('gpurun_out/r2_peer8_$n.json').read().splitlines() if l.startswith('{')]
d=json.loads(lines[-1])
print('N=%d value %.4e ms/step %.4f kernel_ms %.4f e2e %s launches %s graph_steps_rank0 %s dist_verified %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], (d.get('e2e') or {}).get('value'), d['gpu_launches'], d.get('graph_steps_rank0'), d.get('dist_verified')))
print((d.get('e2e') or {}).get('numa_rank0'))
PY
