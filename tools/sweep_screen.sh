for n in 30000 65536 100000; do for q in 1024 4096 16384; do
  echo "exact:"; EBSD_TOPK_SCREEN=0 python tools/topk_once.py $n $q 5
  echo "screen:"; EBSD_TOPK_SCREEN_MIN_ROWS=20000 python tools/topk_once.py $n $q 5
done; done
