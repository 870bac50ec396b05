#!/bin/bash
# Blackwell-native evidence: per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA
# (B200_PROFILING.md "What proves a Blackwell-native kernel").  Usage: bash tools/sass_opcodes.sh > profiles/rNN_sass_opcodes.txt
SO=tml_image_editing_defense_b200/csrc/libtml_b200.so
echo "# cuobjdump -sass $SO  (sm_100a), per kernel: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA load/store,"
echo "# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0)"
cuobjdump -sass $SO | awk '
/Function :/ { fn=$3; order[++n]=fn }
/UTCHMMA\.2CTA/ { c2[fn]++ }
/UTCHMMA/ { mma[fn]++ }
/LDTM/ { ldtm[fn]++ }
/UTMALDG/ { tl[fn]++ }
/UTMASTG/ { ts[fn]++ }
/UTCBAR/ { cb[fn]++ }
/SYNCS/ { sy[fn]++ }
/ HMMA/ { hm[fn]++ }
END { printf "%8s %8s %6s %8s %8s %7s %6s %5s  kernel\n","UTCHMMA","(.2CTA)","LDTM","UTMALDG","UTMASTG","UTCBAR","SYNCS","HMMA";
      for (i=1;i<=n;i++){f=order[i]; printf "%8d %8d %6d %8d %8d %7d %6d %5d  %s\n", mma[f],c2[f],ldtm[f],tl[f],ts[f],cb[f],sy[f],hm[f],f} }' | (read h; echo "$h"; sort -k9 | c++filt 2>/dev/null || cat)
