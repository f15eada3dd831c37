#!/bin/bash
# refresh fast path: tests (fast fail), full suite, then config 4 (d=128) and the headline with the Cholesky fast path on / off
python -m pytest tests/test_emgmm_gpu.py -m gpu -x -q -k "refresh_cholesky or kat or large_shapes" > gpurun_out/pytest_gpu_rf.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_rf.log
tail -5 gpurun_out/pytest_gpu_rf.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_rf_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_rf_full.log
tail -4 gpurun_out/pytest_gpu_rf_full.log
for m in 1 0; do
  MLF_B200_REFRESH_FAST=$m python bench.py --points 2e7 --dims 128 --comps 32 --sigma-c 4 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_cfg4_rf$m.json 2> gpurun_out/bench_cfg4_rf$m.err
  MLF_B200_REFRESH_FAST=$m python bench.py --points 1e6 --dims 2 --comps 8 --sigma-c 6 --steps 50 --warmup 5 --no-e2e --no-cpu > gpurun_out/bench_cfg2_rf$m.json 2> gpurun_out/bench_cfg2_rf$m.err
  MLF_B200_REFRESH_FAST=$m python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_cfg3_rf$m.json 2> gpurun_out/bench_cfg3_rf$m.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_cfg*_rf*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j["roofline"]
        print(f, "ms/step %.3f iter/s %.3f frac %.4f sumLL %.17g"%(j["ms_per_step"], j["em_iter_per_s"], r["frac"], j["sumLL_last"]), {k: round(v,4) for k,v in r["per_iteration_ms"].items()})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-800:])
PY
