#!/bin/bash
# scheduling experiments on the quadratic-form sweep (MLF_B200_QF_EXP): parity tests of the quadratic-form sweep + headline bench under each
for v in ${VALS:-0 1 2 3}; do
  MLF_B200_QF_EXP=$v python -m pytest tests/test_emgmm_gpu.py -m gpu -x -q -k "quadratic or config3" > gpurun_out/pytest_gpu_exp$v.log 2>&1; echo "exp $v pytest exit $?"; tail -2 gpurun_out/pytest_gpu_exp$v.log
  MLF_B200_QF_EXP=$v timeout -s KILL 150 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_exp$v.json 2> gpurun_out/bench_exp$v.err
done
python - <<'PY'
import json, os
for v in os.environ.get("VALS", "0 1 2 3").split():
    f = f"gpurun_out/bench_exp{v}.json"
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j["roofline"]
        print("QF_EXP", v, "ms/step %.2f iter/s %.3f frac %.4f sumLL %.17g"%(j["ms_per_step"], j["em_iter_per_s"], r["frac"], j["sumLL_last"]))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-800:])
PY
