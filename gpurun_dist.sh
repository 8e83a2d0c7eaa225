mkdir -p gpurun_out
N=${1:-2}
for wl in c3 c4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload $wl > gpurun_out/dist_${wl}_$N.json 2>gpurun_out/dist_${wl}_$N.err
echo "rc=$?"; tail -c 1500 gpurun_out/dist_${wl}_$N.err; tail -c 2500 gpurun_out/dist_${wl}_$N.json
done
