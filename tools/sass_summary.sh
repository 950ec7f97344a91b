#!/bin/bash
# Counts the SASS mnemonics that show which Blackwell units the kernels of libflashv.so use (B200_PROFILING.md):
# LDTM/STTM = tcgen05.ld/st (tensor memory), UBLKCP = cp.async.bulk (TMA), SYNCS = mbarrier, FMNMX3 = 3-input
# min/max, UTC*MMA = tcgen05.mma (expected 0: max-plus is not a multiply-accumulate).
LIB=${1:-flash-viterbi_b200/lib/libflashv.so}
TMP=$(mktemp)
cuobjdump -sass "$LIB" > "$TMP"
echo "SASS summary of $LIB ($(date -u +%Y-%m-%dT%H:%MZ), $(cuobjdump --version | tail -1))"
for m in LDTM STTM UBLKCP "SYNCS" FMNMX3 FMNMX "UTC.*MMA" "DADD" "LDS.128" "LDG.E.128" "BAR.SYNC" "ACQBULK\|UTMALDG"; do
  printf "%-14s %6d\n" "$m" "$(grep -c "$m" "$TMP")"
done
echo
echo "instructions per kernel (static SASS):"
awk '/Function :/{name=$3} /^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/{c[name]++} END{for(n in c) print c[n], n}' "$TMP" | sort -rn | head -24
echo
echo "kernels using tensor memory / TMA:"
awk '/Function :/{name=$3} /LDTM|STTM/{t[name]++} /UBLKCP/{u[name]++} END{for(n in t) print "  tcgen05.ld/st:", t[n], n; for(n in u) print "  cp.async.bulk:", u[n], n}' "$TMP"
rm -f "$TMP"
