#!/bin/bash
# Role knock-outs of the screen kernel, compile-time (EXTRA=-DEBSD_SCREEN_KO=n builds in ab_libs/libebsd_ko<n>.so;
# 1 no maximum tree, 2 no TMEM loads, 4 one MMA per half tile instead of three).  Timing only: results are wrong.
N=${1:-1250000}; Q=${2:-80000}
for r in 1 2; do
for l in ebsd_vae_b200/libebsd_b200.so ab_libs/libebsd_ko1.so ab_libs/libebsd_ko3.so ab_libs/libebsd_ko4.so ab_libs/libebsd_ko7.so; do
  echo "$l: $(EBSD_B200_LIB=$l python tools/topk_once.py $N $Q 5 | cut -d, -f1)"
done; done
