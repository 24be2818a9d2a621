#!/bin/bash
# Per-kernel opcode histogram of the shipped library: the Blackwell-native evidence (tcgen05.mma = UTCHMMA / UTCQMMA,
# TMEM loads/stores = LDTM / STTM, TMA = UTMALDG / UTMASTG, tcgen05.commit = UTCBAR, mbarrier = SYNCS), runs without a GPU.
#   tools/sass_histogram.sh > profiles/r02_sass_histogram.txt
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SO="$ROOT/stabletriton_b200/csrc/libstabletriton_b200.so"
echo "# cuobjdump -sass $SO  (sm_100a)"
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn=$3; sub(/^_Z[0-9N]*/, "", fn); names[++n]=$3; cur=$3 }
  /UTCHMMA|UTCQMMA|UTCOMMA/ { c[cur,"UTCxMMA"]++ }
  /LDTM/    { c[cur,"LDTM"]++ }
  /STTM/    { c[cur,"STTM"]++ }
  /UTMALDG/ { c[cur,"UTMALDG"]++ }
  /UTMASTG/ { c[cur,"UTMASTG"]++ }
  /UTCBAR/  { c[cur,"UTCBAR"]++ }
  /SYNCS/   { c[cur,"SYNCS"]++ }
  /MUFU\.EX2/ { c[cur,"MUFU.EX2"]++ }
  /MUFU\.TANH/ { c[cur,"MUFU.TANH"]++ }
  / HMMA| WGMMA/ { c[cur,"legacy_mma"]++ }
  END {
    printf "%8s %6s %6s %8s %8s %7s %6s %9s %10s %10s  %s\n", "UTCxMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MUFU.EX2", "MUFU.TANH", "legacy_mma", "kernel"
    for (i = 1; i <= n; i++) { k = names[i];
      printf "%8d %6d %6d %8d %8d %7d %6d %9d %10d %10d  %s\n", c[k,"UTCxMMA"], c[k,"LDTM"], c[k,"STTM"], c[k,"UTMALDG"], c[k,"UTMASTG"], c[k,"UTCBAR"], c[k,"SYNCS"], c[k,"MUFU.EX2"], c[k,"MUFU.TANH"], c[k,"legacy_mma"], k }
  }' | c++filt 2>/dev/null || true
