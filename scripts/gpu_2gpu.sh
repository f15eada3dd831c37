#!/bin/bash
# 2-rank validation (torchrun, native NCCL): headline mixture and an overlapping one (statistics-first + scatter pass under sharding)
for sc in 8 1; do
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu --sigma-c $sc \
    > gpurun_out/bench_2gpu_sc$sc.json 2> gpurun_out/bench_2gpu_sc$sc.err; echo "2 GPUs sigmaC $sc exit $?"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_2gpu_default.json 2> gpurun_out/bench_2gpu_default.err; echo "2 GPUs default (with e2e) exit $?"
python - <<'PY'
import json
for f in ("bench_2gpu_sc8", "bench_2gpu_sc1", "bench_2gpu_default"):
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); r=j["roofline"]
        print(f, "n_gpus", j["n_gpus"], "ms/step %.2f iter/s %.3f frac %.4f sumLL %.17g"%(j["ms_per_step"], j["em_iter_per_s"], r["frac"], j["sumLL_last"]), r["logits"], j["e2e"] and {k: j["e2e"][k] for k in ("ms_per_step","h2d_gbs","h2d_probe_gbs")})
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-800:])
PY
