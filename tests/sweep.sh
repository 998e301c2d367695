#!/bin/bash
# developer tool (GPU box): bench sweep over launch-geometry overrides: "K B SEG" triples
for cfg in "$@"; do set -- $cfg
  FX8010_TUNE_K=$1 FX8010_TUNE_B=$2 FX8010_TUNE_SEG=$3 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e ${BENCH_ARGS} 2>&1 | python -c "
import sys,json
t=sys.stdin.read().strip().splitlines()
try:
    d=json.loads(t[-1]); print('K=$1 B=$2 SEG=$3', round(d['ms_per_step']*1000,2), 'us frac', round(d['roofline']['frac'],3), d['config']['kernel'])
except Exception as e:
    print('K=$1 B=$2 SEG=$3 FAILED', t[-1][:300] if t else '')
"
done
