#!/bin/bash
# static SASS statistics of one kernel of the built library (dev tool): tools/sass_stats.sh <mangled-name-substring>
LIB=sls_sat_solving_with_deep_learning_b200/lib/libsls_b200.so
FUN=$(cuobjdump -sass $LIB | grep -o "Function : .*$1.*" | head -1 | sed 's/Function : //')
cuobjdump -sass -fun "$FUN" $LIB | grep -E "^\s+/\*[0-9a-f]{4}\*/" > /tmp/sass_fun.txt
echo "$FUN: $(wc -l < /tmp/sass_fun.txt) instructions, BSSY $(grep -c BSSY /tmp/sass_fun.txt), BRA $(grep -c 'BRA' /tmp/sass_fun.txt), IMAD.MOV $(grep -c 'IMAD.MOV' /tmp/sass_fun.txt)"
