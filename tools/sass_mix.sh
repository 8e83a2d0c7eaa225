#!/bin/bash
# tools/sass_mix.sh KERNEL_SUBSTRING [lib]: instruction mix of one kernel of librcs.so (static SASS counts)
lib=${2:-rmf_crowdsim_b200/_lib/librcs.so}
cuobjdump -sass $lib | awk -v k="$1" '/Function :/ {on = index($0, k) > 0} on {print}' > /tmp/_k.sass
echo "lines: $(wc -l < /tmp/_k.sass)"
for op in "LDS" "LDG" " LD\." "STS" "STG" " ST\." "ATOMS" "ATOMG\|RED" "DFMA\|DADD\|DMUL\|DSETP" "FFMA\|FADD\|FMUL\|FSETP" "BAR" "LDL" "STL" "MUFU" "SHFL" "BRA" "UTMA\|UBLKCP\|SYNCS"; do
  printf "%-28s %s\n" "$op" "$(grep -c "$op" /tmp/_k.sass)"
done
