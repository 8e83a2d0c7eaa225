mkdir -p gpurun_out
N=${1:-2}
for wl in c4; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --workload $wl > gpurun_out/dist_${wl}_$N.json 2>gpurun_out/dist_${wl}_$N.err
echo "rc=$?"; grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/dist_${wl}_$N.err | tail -5; python -c "
import json; d=json.loads(open('gpurun_out/dist_${wl}_$N.json').readlines()[-1]); print('$wl N=$N', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'] if d['e2e'] else None)"
done
