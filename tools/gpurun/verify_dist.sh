# NCCL transport parity: N ranks, committed steps, bit-identical to one handle.  gpurun --gpus N -- 'bash tools/gpurun/verify_dist.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --verify-dist --steps ${2:-60} > gpurun_out/verify_dist_$N.json 2> gpurun_out/verify_dist_$N.err
echo "rc=$?"; grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/verify_dist_$N.err | tail -8; cat gpurun_out/verify_dist_$N.json
