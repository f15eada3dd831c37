#!/bin/bash
# strong scaling of the headline job on G GPUs with the final build (resident legs + e2e), as the driver launches it
G=${G:-8}
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $G --steps 20 --warmup 3 --no-cpu \
  > gpurun_out/bench_scale_${G}gpu.json 2> gpurun_out/bench_scale_${G}gpu.err; echo "bench $G GPUs exit $?"
python - <<'PY'
import json, os
G=os.environ.get("G","8")
try:
    j=json.loads(open(f"gpurun_out/bench_scale_{G}gpu.json").read().strip().splitlines()[-1]); r=j["roofline"]
    print("n_gpus", j["n_gpus"], "ms/step %.3f iter/s %.3f value %.4e frac %.4f sumLL %.17g"%(j["ms_per_step"], j["em_iter_per_s"], j["value"], r["frac"], j["sumLL_last"]), {k: round(v,4) for k,v in r["per_iteration_ms"].items()})
    print("  e2e", {k: j["e2e"][k] for k in ("value","ms_per_step","h2d_gbs","h2d_probe_gbs")}, "clocks", j["clocks"])
except Exception as e:
    print("ERR", e, open(f"gpurun_out/bench_scale_{G}gpu.err").read()[-800:])
PY
