# round 2: GPU suite on the current default library, then A/B of library variants ($@) on C3 + C4 (kernel option 3),
# with the round-1 warp kernel (option 2) as the reference line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=3 > gpurun_out/r2_tests.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_tests.log
cp rmf_crowdsim_b200/_lib/librcs.so /tmp/librcs_default.so
run() { # name kernel-option
for wl in c3 c4; do
f=gpurun_out/r2t2_$1_$wl.json
timeout 120 python bench.py --steps 20 --warmup 5 --workload $wl --kernel $2 --skip-e2e --skip-cpu > $f 2>$f.err; tail -c 300 $f.err; python -c "
import json; d=json.loads(open('$f').readlines()[-1]); print('$1 $wl', '%.4e'%d['value'], round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4))"
done
}
for rep in 1 2; do
run warp 2
for v in "$@"; do
cp rmf_crowdsim_b200/_lib/variants/$v.so rmf_crowdsim_b200/_lib/librcs.so
run $v 3
done
cp /tmp/librcs_default.so rmf_crowdsim_b200/_lib/librcs.so
done
