#!/bin/bash
# mid shapes (16 < d <= 32): two-pass k_em<false>+k_mstep vs the large-shape kernels
for dk in "32 32" "24 16" "20 64"; do
set -- $dk
for big in 0 1; do
MLF_B200_BIG=$big python bench.py --points 1e7 --dims $1 --comps $2 --steps 3 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']; print('d=$1 K=$2 BIG=$big ms/step %.2f'%j['ms_per_step'], {k: round(v,2) for k,v in r['per_iteration_ms'].items()}, r['kernel'][:60])"
done
done
