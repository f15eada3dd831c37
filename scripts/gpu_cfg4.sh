#!/bin/bash
python bench.py --points 2e7 --dims 128 --comps 32 --steps 3 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
tail -3 gpurun_out/bench_cfg4.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_cfg4.json").read().strip().splitlines()[-1]); r=j["roofline"]
print("cfg4 ms/step %.2f iter/s %.3f"%(j["ms_per_step"], j["em_iter_per_s"]), r["per_iteration_ms"])
PY
