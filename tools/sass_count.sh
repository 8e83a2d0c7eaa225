#!/bin/bash
# instruction counts per kernel of the built library (quick check that a refactor left the hot kernel alone)
cuobjdump -sass "${1:-rmf_crowdsim_b200/_lib/librcs.so}" 2>/dev/null | awk '/Function : /{f=$3} /^ +\/\*[0-9a-f]+\*\/ /{c[f]++} END{for(k in c) print c[k], k}' | sort -rn | head -${2:-12}
