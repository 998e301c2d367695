#!/bin/bash
# developer tool (GPU box): the stateful configs at larger instance counts (are the kernels HBM-bound once the SMs are full?)
run() { python bench.py --config $1 --instances $2 --steps $3 --warmup 3 --no-cpu-baseline --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', $2, round(d['ms_per_step']*1000,1), 'us frac', round(d['roofline']['frac'],3), d['config']['kernel'])"; }
run cfg4 262144 5; run cfg4 1048576 5; run cfg3 131072 5; run cfg2 65536 20; run cfg5 262144 3
