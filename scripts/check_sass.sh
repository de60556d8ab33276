#!/bin/bash
# Guards against a ptxas 12.9 code-generation bug seen with cp.async + .L2::cache_hint: an ODD
# uniform-register pair as the LDGSTS / LDG / STG descriptor (desc[UR1]) is an illegal instruction
# on sm_100a.  Usage: bash scripts/check_sass.sh   (scans build/obj/*.o; exit 1 if any is found)
bad=0
for f in build/obj/*.o; do
    n=$(cuobjdump -sass "$f" 2>/dev/null | grep -cE "desc\[UR[0-9]*[13579]\]")
    if [ "$n" != "0" ]; then echo "$f: $n instruction(s) with an odd descriptor register pair"; bad=1; fi
done
exit $bad
