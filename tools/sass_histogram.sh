#!/bin/bash
# SASS opcode histogram of the shipped library per kernel: the mnemonics that prove the tcgen05 / TMEM / bulk-copy path
# (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP / UBLKRED = cp.async.bulk / cp.reduce.async.bulk, UTCBAR = tcgen05.commit,
#  SYNCS = mbarrier, FHFMA = mixed-precision FMA of the hi/lo split).   bash tools/sass_histogram.sh > profiles/rXX_sass_opcode_histogram.txt
cd "$(dirname "$0")/.."
LIB=deepphysinet_b200/_lib/libdpn_b200.so
echo "library: $LIB ($(stat -c %s $LIB) bytes), built from $(git rev-parse --short HEAD)"
cuobjdump -sass $LIB | awk '
  /Function :/ { fn=$3; next }
  /^[ \t]+\/\*[0-9a-f]+\*\// {
    op=$2; if (op ~ /^@/) op=$3; sub(/;$/, "", op); split(op, a, "."); base=a[1];
    if (base ~ /^(UTCHMMA|UTCQMMA|LDTM|STTM|UBLKCP|UBLKRED|UBLKPF|UTCBAR|UTCATOMSWS|SYNCS|UTMALDG|FHFMA|F2FP|HMMA|FFMA|DFMA|REDG|ATOMG|ATOMS|STG|LDG|STS|LDS|SHFL|ELECT|UCGABAR_ARV|UCGABAR_WAIT)$/) cnt[fn" "base]++;
    tot[fn]++ }
  END { for (k in cnt) { split(k, b, " "); printf "%s %s %d\n", b[1], b[2], cnt[k] } for (f in tot) printf "%s TOTAL %d\n", f, tot[f] }' |
  sort | c++filt | awk '{ fn=$1; for (i=2;i<=NF-2;i++) fn=fn" "$i; op=$(NF-1); n=$NF; if (fn!=last) { if (last!="") print ""; printf "%s\n   ", fn; last=fn } printf "%s %s | ", op, n } END { print "" }' | cut -c1-400
