#!/bin/bash
# Same-box A/B of the search between builds: tools/ab_topk.sh "LIB_A LIB_B ..." ["N:Q N:Q ..."]
LIBS=$1; SHAPES=${2:-"10000000:10000 1250000:80000 1000000:65536 100000:10000 100000:80000"}
for s in $SHAPES; do
  n=${s%%:*}; q=${s##*:}
  for r in 1 2; do
    for l in $LIBS; do echo "$l $(EBSD_B200_LIB=$l python tools/topk_once.py $n $q 5 | cut -d, -f1)"; done
  done
done
