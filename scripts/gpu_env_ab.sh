#!/bin/bash
# A/B of one environment switch on the headline bench (device-resident legs): ENVVAR=name VALS="a b c"
[ -z "$KEXPR" ] || { python -m pytest tests/test_emgmm_gpu.py -m gpu -x -q -k "$KEXPR" > gpurun_out/pytest_gpu_ab.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_ab.log; tail -4 gpurun_out/pytest_gpu_ab.log; }
for v in $VALS; do
  env $ENVVAR=$v timeout -s KILL ${BTO:-150} python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu $BARGS > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
done
python - <<'PY'
import json, glob, os
for v in os.environ["VALS"].split():
    f = f"gpurun_out/bench_ab_{v}.json"
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j["roofline"]
        print(os.environ["ENVVAR"], v, "ms/step %.2f iter/s %.3f frac %.4f sumLL %.17g"%(j["ms_per_step"], j["em_iter_per_s"], r["frac"], j["sumLL_last"]), r["per_iteration_ms"], r.get("logits"))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-800:])
PY
