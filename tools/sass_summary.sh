#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMA / mma.sync use (B200_PROFILING.md), from the built library.
so=lip2speech_b200/lib/libl2s_b200.so
cuobjdump -sass $so | awk '
/Function : / { name=$3 }
/UTCHMMA/ { a[name]++ } /UTMALDG/ { b[name]++ } /LDTM/ { c[name]++ } /UBLKCP/ { d[name]++ } /UTCBAR/ { e[name]++ } /HMMA\.1688\.F32\.TF32/ { f[name]++ }
/STRONG\.GPU/ { g[name]++ } /SYNCS/ { h[name]++ } /LDGSTS/ { i[name]++ }
END { printf "%-90s %8s %8s %6s %7s %7s %10s %11s %6s %7s\n", "kernel (mangled)", "UTCHMMA", "UTMALDG", "LDTM", "UBLKCP", "UTCBAR", "HMMA.TF32", "STRONG.GPU", "SYNCS", "LDGSTS";
  for (k in a) n[k]=1; for (k in b) n[k]=1; for (k in c) n[k]=1; for (k in d) n[k]=1; for (k in e) n[k]=1; for (k in f) n[k]=1; for (k in g) n[k]=1; for (k in i) n[k]=1;
  for (k in n) printf "%-90s %8d %8d %6d %7d %7d %10d %11d %6d %7d\n", substr(k,1,90), a[k], b[k], c[k], d[k], e[k], f[k], g[k], h[k], i[k] }' | sort
