#!/bin/bash
# Same-box A/B of programmatic dependent launch in the encoder chain (EBSD_ENCODER_PDL=0 serialises the launches):
#   tools/ab_pdl.sh [ROUNDS]      -> encoder alone on 10 000 patterns, then the whole bench step
R=${1:-3}
for r in $(seq $R); do for v in 1 0; do
  echo -n "encoder PDL=$v  "; EBSD_ENCODER_PDL=$v timeout 120 python tools/encode_once.py 10000 4 8
done; done
for r in 1 2; do for v in 1 0; do
  EBSD_ENCODER_PDL=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sweeps 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench PDL=$v', 'value %.1f k  e2e %.1f k  enc %.2f ms  step %.2f ms' % (d['value']/1e3, d['e2e']['value']/1e3, d['stages']['encoder_ms'], d['ms_per_step']), d['clocks'])"
done; done
