mkdir -p gpurun_out
N=${1:-2}
RCS_E2E_TRACE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/dist_tr_$N.json 2>gpurun_out/dist_tr_$N.err
echo "rc=$?"; grep "e2e host" gpurun_out/dist_tr_$N.err; python -c "
import json; d=json.loads(open('gpurun_out/dist_tr_$N.json').readlines()[-1]); print('N=$N', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
