#!/bin/bash
# Same-box A/B of the whole bench step between library builds: tools/ab_bench.sh "LIB_A LIB_B ..." [ROUNDS]
LIBS=$1; R=${2:-2}
for r in $(seq $R); do for l in $LIBS; do
  EBSD_B200_LIB=$l python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sweeps 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$l', 'value %.1f k  e2e %.1f k  enc %.2f ms' % (d['value']/1e3, d['e2e']['value']/1e3, d['stages']['encoder_ms']), {k:v['us_per_1184_patterns'] for k,v in d['roofline']['blocks'].items()})"
done; done
