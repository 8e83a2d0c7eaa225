# round 2, call 2: GPU suite on the tile kernel (default), then warp (L1) vs tile (smem) vs tile (bulk copies), C3 + C4
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2_tests.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_tests.log
cp rmf_crowdsim_b200/_lib/librcs.so /tmp/librcs_default.so
run() { # name kernel-option
for wl in c3 c4; do
f=gpurun_out/r2t1_$1_$wl.json
timeout 120 python bench.py --steps 20 --warmup 5 --workload $wl --kernel $2 --skip-e2e --skip-cpu > $f 2>$f.err; tail -c 300 $f.err; python -c "
import json; d=json.loads(open('$f').readlines()[-1]); print('$1 $wl', '%.4e'%d['value'], round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4))"
done
}
for rep in 1 2; do
run warp 2
run tile 3
cp rmf_crowdsim_b200/_lib/variants/tile_tma.so rmf_crowdsim_b200/_lib/librcs.so
run tile_tma 3
cp /tmp/librcs_default.so rmf_crowdsim_b200/_lib/librcs.so
done
