#!/bin/bash
# N-GPU scaling run: torchrun as the driver launches it
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_g${N}.json 2> gpurun_out/bench_g${N}.err
tail -3 gpurun_out/bench_g${N}.err
python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_g${N}.json").read().strip().splitlines()[-1]); r=j["roofline"]
    print("N=${N}", "ms/step %.2f iter/s %.3f value %.3e frac %.3f"%(j["ms_per_step"], j["em_iter_per_s"], j["value"], r["frac"]), r["per_iteration_ms"], "e2e", j["e2e"])
except Exception as e:
    print("ERR", e)
PY
