#!/bin/bash
# Blackwell-native evidence from the shipped library: per kernel, how many tcgen05 MMA (UTCHMMA = kind::f16/tf32, UTCQMMA = kind::f8f6f4),
# TMEM load/store (LDTM/STTM), bulk copy (UBLKCP) and TMA tensor-map (UTMALDG) instructions its SASS holds.
#   bash tools/sass_summary.sh > profiles/rNN_sass_opcount_summary.txt
LIB=${1:-mirror_nerf_b200/lib/libmnrf.so}
echo "cuobjdump -sass $LIB  (sm_100a)   count mnemonic kernel"
cuobjdump -sass "$LIB" 2>/dev/null | grep -E "Function :|UTC[A-Z]*MMA|LDTM|STTM|UBLKCP|UTMALDG|SYNCS" \
 | awk '/Function/{fn=$3} /UTC[A-Z]*MMA|LDTM|STTM|UBLKCP|UTMALDG/{for(i=1;i<=NF;i++) if($i ~ /^(UTC[A-Z]*MMA|LDTM|STTM|UBLKCP|UTMALDG)/){split($i,a,"."); c[fn" "a[1]]++}} END{for(k in c) print c[k], k}' \
 | sort -k2,2 -k3,3 | awk '{printf "%6d %-10s %s\n",$1,$3,$2}' | c++filt 2>/dev/null | sed 's/mnrf::(anonymous namespace):://'
