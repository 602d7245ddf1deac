#!/bin/bash
# mnemonic counts of the shipped library (Blackwell-native evidence per /opt/skills/guides/B200_PROFILING.md):
#   scripts/sass_mnemonics.sh > profiles/r02_sass_mnemonics.txt
LIB=rwkvtts_b200/librwkvtts_wkv7.so
SASS=$(mktemp); cuobjdump -sass $LIB > $SASS
echo "# cuobjdump -sass $LIB | mnemonic counts (commit $(git rev-parse --short HEAD)); Blackwell-native evidence per /opt/skills/guides/B200_PROFILING.md"
echo "# tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR, cp.async.bulk -> UBLKCP, cp.async.bulk.tensor -> UTMALDG, cp.async -> LDGSTS, mma.sync -> HMMA, multimem.* -> LDGMC / REDG-type"
for m in UTCHMMA UTCQMMA LDTM STTM UTCBAR UBLKCP UTMALDG UTMASTG UTMAPF LDGSTS HMMA SYNCS; do printf "%-10s %s\n" $m $(grep -c "[ .]$m" $SASS); done
echo; echo "# per kernel (chunked WKV-7 kernels)"
awk '/Function : /{fn=$3} /UTCHMMA/{a[fn]++} /LDTM/{b[fn]++} /STTM/{c[fn]++} /UBLKCP/{d[fn]++} /UTMALDG/{e[fn]++} /LDGSTS/{f[fn]++} /HMMA/{g[fn]++} END{for (k in a) printf "%s UTCHMMA=%d LDTM=%d STTM=%d UBLKCP=%d UTMALDG=%d LDGSTS=%d HMMA=%d\n", k, a[k], b[k], c[k], d[k], e[k], f[k], g[k]}' $SASS | sort
echo; echo "# tensor-map copies of the backward's stage A (one per input tensor and chunk), first occurrences"
grep -m 4 "UTMALDG" $SASS | sed 's/^ *//' | cut -c1-120
echo; echo "# fused ZeRO exchange (csrc/adam.cu adam_p2p_kernel<true>): multimem.ld_reduce / multimem.st through the NVSwitch"
grep -E "LDGMC|MULTIMEM|\.MMLD|STGMC|REDGMC" $SASS | sed 's/^ *//' | awk '{print $2}' | sort | uniq -c | head
rm -f $SASS
