#!/bin/bash
# full GPU suite, then the headline bench (device-resident legs only) under each k_em16 variant (MLF_B200_EM16_VAR)
[ -n "$SKIP_PYTEST" ] || { python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_v.log; }
[ -n "$SKIP_PYTEST" ] || tail -5 gpurun_out/pytest_gpu_v.log
for v in ${VARS:-0 1 2 3}; do
  MLF_B200_EM16_VAR=$v timeout -s KILL ${BTO:-120} python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_var$v.json 2> gpurun_out/bench_var$v.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_var*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j["roofline"]
        print(f, "ms/step %.2f iter/s %.3f frac %.4f sumLL %.17g"%(j["ms_per_step"], j["em_iter_per_s"], r["frac"], j["sumLL_last"]), r["per_iteration_ms"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-800:])
PY
