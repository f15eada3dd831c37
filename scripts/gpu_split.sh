#!/bin/bash
# statistics split over warps at small K: full suite, then config 2 and a K = 16 shape with and without it
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_split.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_split.log
tail -3 gpurun_out/pytest_gpu_split.log
for v in 1 0; do
  MLF_B200_EM16_SPLIT=$v python bench.py --points 1e6 --dims 2 --comps 8 --sigma-c 6 --steps 50 --warmup 5 --no-e2e --no-cpu > gpurun_out/bench_split_cfg2_$v.json 2> gpurun_out/bench_split_cfg2_$v.err
  MLF_B200_EM16_SPLIT=$v python bench.py --points 2e7 --dims 16 --comps 16 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_split_k16_$v.json 2> gpurun_out/bench_split_k16_$v.err
  MLF_B200_EM16_SPLIT=$v python bench.py --points 2e7 --dims 8 --comps 32 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_split_k32_$v.json 2> gpurun_out/bench_split_k32_$v.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_split_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j["roofline"]
        print(f, "ms/step %.4f iter/s %.2f frac %.4f sumLL %.17g"%(j["ms_per_step"], j["em_iter_per_s"], r["frac"], j["sumLL_last"]))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-600:])
PY
