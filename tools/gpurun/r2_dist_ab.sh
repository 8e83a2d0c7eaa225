# graphs on / off on N GPUs: gpurun --gpus N -- 'bash tools/gpurun/r2_dist_ab.sh N'
n=${1:-8}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
for g in 0 1 0 1; do
RCS_GRAPHS=$g timeout 300 python bench.py --gpus $n --steps 20 --warmup 5 --skip-e2e --skip-verify > gpurun_out/r2_dab_${n}_$g.json 2> gpurun_out/r2_dab_${n}_$g.err; echo "rc=$?"
python - <<PY
import json
lines=[l for l in open('gpurun_out/r2_dab_${n}_$g.json').read().splitlines() if l.startswith('{')]
d=json.loads(lines[-1])
print('graphs=$g N=%d value %.4e ms/step %.4f kernel_ms %.4f launches %s graph_steps_rank0 %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches'], d.get('graph_steps_rank0')))
PY
done
