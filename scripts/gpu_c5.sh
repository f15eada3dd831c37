#!/bin/bash
# diagonal sweep: tests, then the config-5 shard with the expanded-form P1 (default) and the lane = component P1
python -m pytest tests/test_emgmm_gpu.py -m gpu -x -q -k "diag or config5" > gpurun_out/pytest_gpu_c5.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_c5.log
tail -12 gpurun_out/pytest_gpu_c5.log
for q in ${QFS:-1 0}; do
  MLF_B200_QF=$q python bench.py --points 1.25e8 --dims 8 --comps 256 --cov-type diag --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_cfg5_qf$q.json 2> gpurun_out/bench_cfg5_qf$q.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_cfg5_qf*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j["roofline"]
        print(f, "ms/step %.2f iter/s %.3f value %.3e frac %.4f sumLL %.17g"%(j["ms_per_step"], j["em_iter_per_s"], j["value"], r["frac"], j["sumLL_last"]), {k: round(v,3) for k,v in r["per_iteration_ms"].items()}, r.get("logits"))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-800:])
PY
