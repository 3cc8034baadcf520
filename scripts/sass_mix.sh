#!/bin/bash
# static SASS opcode histogram of one kernel in the built library: scripts/sass_mix.sh <substring-of-mangled-name>
LIB=cusumtools_b200/libcusumtools_b200.so
FN=$(cuobjdump -elf $LIB 2>/dev/null | grep -o "_Z[A-Za-z0-9_]*$1[A-Za-z0-9_]*" | sort -u | head -1)
echo "kernel: $FN"
cuobjdump -sass -fun "$FN" $LIB | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed -E 's/\..*//;s/;//' | sort | uniq -c | sort -rn | head -${2:-16}
