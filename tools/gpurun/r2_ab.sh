# quick A/B of library variants on C4 (+C3), kernel option 3, no test suite: bash tools/gpurun/r2_ab.sh v1 v2 ...
mkdir -p gpurun_out
cp rmf_crowdsim_b200/_lib/librcs.so /tmp/librcs_default.so
for rep in 1 2; do
for v in "$@"; do
cp rmf_crowdsim_b200/_lib/variants/$v.so rmf_crowdsim_b200/_lib/librcs.so
for wl in c3 c4; do
f=gpurun_out/r2ab_${v}_$wl.json
timeout 120 python bench.py --steps 20 --warmup 5 --workload $wl --skip-e2e --skip-cpu --skip-secondary > $f 2>$f.err; tail -c 300 $f.err; python -c "
import json; d=json.loads(open('$f').readlines()[-1]); print('$v $wl', '%.4e'%d['value'], round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4))"
done
done
done
cp /tmp/librcs_default.so rmf_crowdsim_b200/_lib/librcs.so
