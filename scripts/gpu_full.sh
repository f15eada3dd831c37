#!/bin/bash
# the driver's round-end sequence on one GPU: full GPU suite, smoke, default bench (with e2e / cpu / parity legs), reference arm
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_full.log
tail -4 gpurun_out/pytest_gpu_full.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1]); r=j["roofline"]
print("ms/step %.2f iter/s %.3f frac %.4f"%(j["ms_per_step"], j["em_iter_per_s"], r["frac"]), r["logits"])
print("e2e", j["e2e"] and {k: j["e2e"][k] for k in ("value","ms_per_step","h2d_gbs")})
print("cpu", j["cpu_baseline"])
print("parity", j["parity_check"])
print("clocks", j["clocks"], "launches", j["gpu_launches"])
print(open("gpurun_out/bench_ref.json").read()[-600:])
PY
