#!/bin/bash
# Per-kernel counts of the Blackwell-specific SASS instructions in libvoicemap_b200.so (machine-checkable evidence that
# the hot kernels run on tcgen05 tensor cores with TMEM accumulators and TMA):  UTCHMMA / UTCQMMA = tcgen05.mma
# kind::f16 / kind::f8f6f4, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA bulk tensor load / store, UTCBAR = tcgen05.commit,
# SYNCS = mbarrier operations.   tools/sass_counts.sh > profiles/r02_sass_counts.txt
LIB=${1:-voicemap_b200/libvoicemap_b200.so}
echo "# $(date -u +%Y-%m-%dT%H:%MZ)  $(basename $LIB)  $(stat -c %s $LIB) bytes  cuobjdump -sass, sm_100a"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3; order[++n]=fn }
  /UTCHMMA/ { c[fn,"UTCHMMA"]++ } /UTCQMMA/ { c[fn,"UTCQMMA"]++ } /LDTM/ { c[fn,"LDTM"]++ }
  /UTMALDG/ { c[fn,"UTMALDG"]++ } /UTMASTG/ { c[fn,"UTMASTG"]++ } /UTCBAR/ { c[fn,"UTCBAR"]++ } /SYNCS/ { c[fn,"SYNCS"]++ }
  /^[ \t]*\/\*[0-9a-f]+\*\// { ins[fn]++ }
  END {
    printf "%8s %8s %6s %8s %8s %7s %6s %7s  %s\n", "UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "instrs", "kernel"
    for (i = 1; i <= n; i++) { f = order[i];
      printf "%8d %8d %6d %8d %8d %7d %6d %7d  %s\n", c[f,"UTCHMMA"], c[f,"UTCQMMA"], c[f,"LDTM"], c[f,"UTMALDG"], c[f,"UTMASTG"], c[f,"UTCBAR"], c[f,"SYNCS"], ins[f], f } }' | c++filt | sed -E 's/\((CUtensorMap_st|float|double|unsigned|int|void|vm::|__half|long|char).*$//'
