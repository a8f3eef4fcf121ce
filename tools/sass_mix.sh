#!/bin/bash
# usage: tools/sass_mix.sh <mangled-or-substring kernel name> [top N]   -- opcode histogram of one kernel of checkm_b200/libckm.so
# (how the `.RELU` / zero-register quirk of round 1 was found: 587 PRMT in ssv_kernel<32> where the source implied ~30)
set -e
LIB="$(dirname "$0")/../checkm_b200/libckm.so"
FN=$(cuobjdump -elf "$LIB" 2>/dev/null | grep -o "_ZN3ckm[A-Za-z0-9_]*$1[A-Za-z0-9_]*" | sort -u | head -1)
[ -z "$FN" ] && { echo "no kernel matching $1"; exit 1; }
echo "== $FN"
cuobjdump -sass -fun "$FN" "$LIB" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${2:-16}
