# peer-store transport on N GPUs: normal path, then the collective fallback to NCCL when one rank reports a failed mapping
n=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
for f in none 1; do
  [ "$f" = none ] && unset RCS_PEER_FAIL_RANK || export RCS_PEER_FAIL_RANK=$f
  timeout 400 python bench.py --gpus $n --steps 20 --warmup 5 --skip-e2e > gpurun_out/r2_pfb_$f.json 2> gpurun_out/r2_pfb_$f.err; echo "bench fail_rank=$f rc=$?"
  grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/r2_pfb_$f.err | tail -c 600
  python - <<PY
import json
lines=[l for l in open('gpurun_out/r2_pfb_$f.json').read().splitlines() if l.startswith('{')]
d=json.loads(lines[-1])
print('N=%d value %.4e ms/step %.4f kernel_ms %.4f dist_verified %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d.get('dist_verified')))
print(d['config']['parallelism'][:110])
PY
done
