#!/bin/bash
# per-GPU shard of BASELINE config 5: N=1e9/8 points, d=8, K=256, diagonal covariances (extension)
python bench.py --points 1.25e8 --dims 8 --comps 256 --cov-type diag --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
tail -3 gpurun_out/bench_cfg5.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_cfg5.json").read().strip().splitlines()[-1]); r=j["roofline"]
print("cfg5 shard ms/step %.2f iter/s %.3f value %.3e"%(j["ms_per_step"], j["em_iter_per_s"], j["value"]), r["per_iteration_ms"])
PY
