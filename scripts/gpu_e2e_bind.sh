#!/bin/bash
# does binding the single rank to the GPU-local cores change the H2D rate of the e2e leg?
for nb in 0 1; do
MLF_BENCH_NO_BIND=$nb python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=j['e2e']; print('NO_BIND=$nb', 'ms/step %.2f'%j['ms_per_step'], 'e2e ms %.2f h2d %.2f GB/s'%(e['ms_per_step'], e['h2d_gbs']), e['host_affinity'])"
done
nvidia-smi topo -m 2>/dev/null | head -14
