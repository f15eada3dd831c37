#!/bin/bash
# quadratic-form sweep: its tests first (fast fail), then the headline bench with the quadratic form (default) and with whitening only
python -m pytest tests/test_emgmm_gpu.py -m gpu -x -q -k "${KEXPR:-quadratic or em16_scheduling or config3 or em_iterations}" > gpurun_out/pytest_gpu_qf.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_qf.log
tail -15 gpurun_out/pytest_gpu_qf.log
for q in ${QFS:-1 0}; do
  MLF_B200_QF=$q timeout -s KILL ${BTO:-150} python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_qf$q.json 2> gpurun_out/bench_qf$q.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_qf*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j["roofline"]
        print(f, "ms/step %.2f iter/s %.3f frac %.4f sumLL %.17g"%(j["ms_per_step"], j["em_iter_per_s"], r["frac"], j["sumLL_last"]), r["per_iteration_ms"], r.get("logits"))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-800:])
PY
