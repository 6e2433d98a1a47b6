#!/bin/bash
# SASS opcode histogram of the in-tree library: the tcgen05 / TMEM / TMA mnemonics that prove the kernels are Blackwell-native
# (B200_PROFILING.md: UTCHMMA / UTCQMMA = tcgen05.mma, UTMALDG = TMA tile load, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit).
SO=${1:-ipercore_b200/libiper_b200.so}
OUT=${2:-profiles/r02_sass_histogram.txt}
{
  echo "# cuobjdump -sass $SO  (sm_100a cubins), $(date -u +%Y-%m-%dT%H:%MZ)"
  echo "# per-kernel counts of the Blackwell tensor / TMA opcodes, then the library-wide totals"
  cuobjdump -sass "$SO" | awk '
    /Function :/ { fn=$3 }
    { for (i=1;i<=NF;i++) if ($i ~ /^(UTC[A-Z0-9.]+|UTMALDG[A-Z0-9.]*|UTMAPF[A-Z0-9.]*|LDTM[A-Za-z0-9.]*|STTM[A-Za-z0-9.]*|SYNCS[A-Z0-9.]*|UTCATOMSWS[A-Z0-9.]*|HMMA[A-Z0-9.]*|FFMA|DADD|ATOMS[A-Z0-9.]*|REDG[A-Z0-9.]*|RED[A-Z0-9.]*)$/) { c[fn" "$i]++; t[$i]++ } }
    END { for (k in c) print c[k], k | "sort -k2,2 -k1,1nr"; close("sort -k2,2 -k1,1nr"); print "# ---- totals ----"; for (k in t) print t[k], k | "sort -k1,1nr" }'
} > "$OUT"
echo "wrote $OUT"
