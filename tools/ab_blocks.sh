#!/bin/bash
# Same-box A/B of encoder blocks between two builds of the library: tools/ab_blocks.sh LIB_A LIB_B "LAYERS" [NIMG] [ROUNDS]
# (box-to-box differences between gpurun calls are 3-7 %, larger than most single changes)
A=$1; B=$2; LAYERS=${3:-"1 2 3 4 5 6 7 8 9"}; N=${4:-1184}; R=${5:-2}
for r in $(seq $R); do
  for l in $LAYERS; do
    a=$(EBSD_B200_LIB=$A python tools/time_block.py $l $N 20 | sed 's/.*: *\([0-9.]*\) us.*/\1/')
    b=$(EBSD_B200_LIB=$B python tools/time_block.py $l $N 20 | sed 's/.*: *\([0-9.]*\) us.*/\1/')
    echo "round $r layer $l: A $a us  B $b us"
  done
done
