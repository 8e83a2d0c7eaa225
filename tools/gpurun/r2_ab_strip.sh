# A/B of library variants on a plain handle and on a strip group of one rank (C4): bash tools/gpurun/r2_ab_strip.sh v1 v2 ...
mkdir -p gpurun_out
cp rmf_crowdsim_b200/_lib/librcs.so /tmp/librcs_default.so
for rep in 1 2; do
for v in "$@"; do
cp rmf_crowdsim_b200/_lib/variants/$v.so rmf_crowdsim_b200/_lib/librcs.so
timeout 300 python tools/strip_overhead.py > gpurun_out/r2abs_${v}_$rep.log 2>&1; echo "$v rc=$?"; tail -2 gpurun_out/r2abs_${v}_$rep.log
done
done
cp /tmp/librcs_default.so rmf_crowdsim_b200/_lib/librcs.so
