#!/bin/bash
# fixed per-iteration cost: the headline shape at one eighth of the points on one GPU (what each of 8 ranks sees, minus the all-reduce)
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_lat.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_lat.log
tail -3 gpurun_out/pytest_gpu_lat.log
python bench.py --points 1.25e7 --steps 40 --warmup 5 --no-e2e --no-cpu > gpurun_out/bench_lat.json 2> gpurun_out/bench_lat.err
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_lat_full.json 2> gpurun_out/bench_lat_full.err
python - <<'PY'
import json
for f in ("bench_lat", "bench_lat_full"):
    j=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); r=j["roofline"]
    print(f, "ms/step %.4f frac %.4f"%(j["ms_per_step"], r["frac"]), {k: round(v,4) for k,v in r["per_iteration_ms"].items()})
PY
