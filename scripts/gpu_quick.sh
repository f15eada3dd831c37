#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_q.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_q.log
tail -5 gpurun_out/pytest_gpu_q.log
python bench.py --points 4e6 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_q_small.json 2> gpurun_out/bench_q_small.err
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_q_full.json 2> gpurun_out/bench_q_full.err
python - <<'PY'
import json
for f in ["bench_q_small","bench_q_full"]:
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); r=j["roofline"]
        print(f, "ms/step %.2f iter/s %.3f frac %.3f"%(j["ms_per_step"], j["em_iter_per_s"], r["frac"]), r["per_iteration_ms"])
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-500:])
PY
