# peer-store halo transport against ncclSend / ncclRecv on N GPUs: verify-dist + bench, both transports
n=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
for t in peer nccl; do
  export RCS_HALO=$t
  if [ "$t" = peer ] || [ "$n" = 2 ]; then
    timeout 300 python bench.py --gpus $n --verify-dist > gpurun_out/r2_peer_verify_${n}_$t.json 2> gpurun_out/r2_peer_verify_${n}_$t.err; echo "verify $t rc=$?"
    tail -c 600 gpurun_out/r2_peer_verify_${n}_$t.err; tail -c 1200 gpurun_out/r2_peer_verify_${n}_$t.json
  fi
  x=""; [ "$t" = nccl ] && [ "$n" != 2 ] && x="--skip-e2e"
  timeout 600 python bench.py --gpus $n --steps 20 --warmup 5 $x > gpurun_out/r2_peer_${n}_$t.json 2> gpurun_out/r2_peer_${n}_$t.err; echo "bench $t rc=$?"
  tail -c 800 gpurun_out/r2_peer_${n}_$t.err
  python - <<PY
import json
lines=[l for l in open('gpurun_out/r2_peer_${n}_$t.json').read().splitlines() if l.startswith('{')]
d=json.loads(lines[-1])
print('$t N=%d value %.4e ms/step %.4f kernel_ms %.4f e2e %s launches %s graph_steps_rank0 %s dist_verified %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], (d.get('e2e') or {}).get('value'), d['gpu_launches'], d.get('graph_steps_rank0'), d.get('dist_verified')))
print(d['config']['parallelism'])
PY
done
