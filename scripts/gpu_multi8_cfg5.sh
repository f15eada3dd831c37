#!/bin/bash
# config 5 on 8 GPUs only (N=1e9, d=8, K=256, diagonal)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --points 1e9 --dims 8 --comps 256 --cov-type diag --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_g8_cfg5.json 2> gpurun_out/bench_g8_cfg5.err
tail -2 gpurun_out/bench_g8_cfg5.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_g8_cfg5.json").read().strip().splitlines()[-1]); r=j["roofline"]
print("cfg5 N=1e9 8 GPUs ms/step %.2f iter/s %.3f value %.3e frac %.3f"%(j["ms_per_step"], j["em_iter_per_s"], j["value"], r["frac"]), r["per_iteration_ms"])
PY
