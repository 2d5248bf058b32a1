#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md):
#   tools/sass_summary.sh [lib] > profiles/rNN_sass_summary.txt
LIB=${1:-ebsd_vae_b200/libebsd_b200.so}
echo "cuobjdump -sass $LIB (sm_100a) -- per kernel: tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM, TMA = UTMALDG / UTMASTG / UTMAPF (prefetch) / UBLKCP (bulk copy), legacy tensor path = HMMA"
printf "%-100s %8s %6s %8s %8s %7s %7s %6s %6s\n" kernel UTCxMMA LDTM UTMALDG UTMASTG UTMAPF UBLKCP HMMA lines
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function :/ { if (name != "") out(); name=$3; mma=ld=tl=ts=pf=bc=hm=n=0; next }
/UTC[A-Z]*MMA/ {mma++} /LDTM/ {ld++} /UTMALDG/ {tl++} /UTMASTG/ {ts++} /UTMAPF/ {pf++} /UBLKCP/ {bc++} / HMMA/ {hm++}
/^ +\/\*[0-9a-f]{4,}\*\// {n++}
function out() { printf "%-100s %8d %6d %8d %8d %7d %7d %6d %6d\n", substr(name,1,100), mma, ld, tl, ts, pf, bc, hm, n }
END { out() }' | while read -r line; do
  name=$(echo "$line" | awk '{print $1}'); dem=$(echo "$name" | c++filt | sed 's/(.*//' | cut -c1-100)
  echo "$line" | awk -v d="$dem" '{printf "%-100s %8s %6s %8s %8s %7s %7s %6s %6s\n", d, $2,$3,$4,$5,$6,$7,$8,$9}'
done | sort
