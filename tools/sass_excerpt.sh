#!/bin/bash
# SASS evidence for profiles/: per kernel, how often the mnemonics that prove bulk-TMA staging (UBLKCP + SYNCS mbarrier),
# asynchronous copies (LDGSTS), 128-bit CAS and vector float reductions occur; and that no tensor-core / TMEM instruction
# exists (nothing on this path is a contraction).  usage: bash tools/sass_excerpt.sh > profiles/rNN_sass_excerpt.txt
SO=rovinasemanticsegmentation_b200/librss.so
echo "# cuobjdump -sass $SO | grep -E 'UBLKCP|SYNCS|LDGSTS|ATOMG.*128|REDG|ATOMS|UTMA|UTC.*MMA|LDTM|STTM' (count per kernel and mnemonic)"
cuobjdump -lelf $SO | head -12
cuobjdump -sass $SO | awk '
/Function :/ { f=$3; next }
{ for (i=1;i<=NF;i++) if ($i ~ /^(UBLKCP|SYNCS|LDGSTS|ATOMG|REDG|ATOMS|UTMA|UTC|LDTM|STTM|ARRIVES|MEMBAR|RED\.)/) { split($i,a,";"); c[a[1]" "f]++; break } }
END { for (k in c) print c[k], k }' | c++filt | sed 's/(.*//' | grep -v -E 'kernel<[12346],|kernel<[0-9], [0-9], 0' | sort -k3,3 -k2,2 | awk '{n=$1; m=$2; $1=""; $2=""; printf "%5d  %-42s %s\n", n, m, $0}'
echo "# tensor-core / TMEM / tensor-map TMA instructions:"
cuobjdump -sass $SO | grep -c -E "UTCMMA|UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG|HMMA|IMMA" || true
