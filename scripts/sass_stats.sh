#!/usr/bin/env bash
# static SASS statistics of one scan_kernel instantiation: usage sass_stats.sh <mangled-substring> [object]
OBJ=${2:-toppra_b200/csrc/tb_scan.o}
cuobjdump -sass $OBJ | awk -v pat="$1" '
/Function :/ {f = (index($0, pat) > 0)}
f && /^ +\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f][0-9a-f]?\*\// {n++; if ($0 ~ /S2R/) s2r++; if ($0 ~ /LDL/) ldl++; if ($0 ~ /STL/) stl++; if ($0 ~ /LDC/) ldc++; if ($0 ~ /REDUX/) redux++; if ($0 ~ /SHFL/) shfl++; if ($0 ~ /MUFU.RCP64H/) rcp++; if ($0 ~ /BSSY|BSYNC/) bs++}
END {printf "instr %d  S2R %d  LDL %d  STL %d  LDC %d  REDUX %d  SHFL %d  RCP64H %d  BSSY/BSYNC %d\n", n, s2r, ldl, stl, ldc, redux, shfl, rcp, bs}'
